// libgitb200.so -- C ABI + host-side engine of the B200-native GIT captioning hot path.
// See include/gitb200.h for the contract of every entry point and the reference function it replaces.
//
// Device data layout (all engine-owned, HBM resident):
//   weights      : GEMM operands bf16 [out, in] (the nn.Linear layout is already K-major), biases /
//                  LayerNorm / embeddings fp32; decoder q,k,v fused to one [2304, 768] matrix; patch kernel
//                  flattened to [d, 3*p*p] zero-padded to a multiple of 64 columns.
//   encoder      : residual stream x fp32 [NI*L, d]; GEMM A operands (LN output, attention context, MLP
//                  hidden) bf16 row-major; packed qkv bf16 [NI*L, 3d].
//   image K/V    : bf16 [layer][k|v][B][M][768]  (token-major rows: a decode-step reader streams contiguous
//                  1536-byte rows; written once by the prefill QKV GEMM epilogue, shared by all beams).
//   text K/V     : bf16 [layer][k|v][rows][T_alloc][768] + int32 src_row[rows][T_alloc] indirection for beams.
//   decode step  : fp32 row state [rows, 768], bf16 copy for the GEMMs, fp32 qkv / logits.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <set>
#include <string>
#include <utility>
#include <vector>

#include "../../include/gitb200.h"
#include "attention.cuh"
#include "constrained.cuh"
#include "decode_mega.cuh"
#include "gemm.cuh"
#include "gemm2.cuh"
#include "preproc.cuh"
#include "ptx.cuh"
#include "rowops.cuh"
#include "search.cuh"

using namespace gitb200;
typedef __nv_bfloat16 bf16;

#define GITB200_ABI_VERSION 5

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_create_error;

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  bool owned = true;   // false: borrowed from another engine (gitb200_share_weights)
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (!owned) return cudaErrorInvalidValue;   // a borrowed buffer is never re-allocated
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaSuccess) cap = bytes;
    return e;
  }
  void release() {
    if (p && owned) cudaFree(p);
    p = nullptr;
    cap = 0;
    owned = true;
  }
  void borrow(const DevBuf& o) {
    release();
    p = o.p;
    cap = o.cap;
    owned = false;
  }
  template <typename T>
  T* as() const { return reinterpret_cast<T*>(p); }
};

struct EncLayer {
  DevBuf wqkv, bqkv, wo, bo, ln1g, ln1b, ln2g, ln2b, w1, b1, w2, b2;
};
struct DecLayer {
  DevBuf wqkv, bqkv, wo, bo, lnag, lnab, w1, b1, w2, b2, lnog, lnob;
  DevBuf m_wqkv, m_wo, m_w1, m_w2;   // fragment-packed 8-feature tiles for decode_mega_kernel (built by finalize_weights)
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct TmapKey {
  const void* ptr;
  long long rows, cols, ld;
  int box_rows;   // negative: un-swizzled box (decode attention K/V slices)
  bool operator<(const TmapKey& o) const {
    if (ptr != o.ptr) return ptr < o.ptr;
    if (rows != o.rows) return rows < o.rows;
    if (cols != o.cols) return cols < o.cols;
    if (ld != o.ld) return ld < o.ld;
    return box_rows < o.box_rows;
  }
};

struct gitb200_engine {
  gitb200_config cfg;
  int device = 0;
  int num_sms = 148;
  std::string err;
  int64_t launches = 0;
  bool use_graph = true;
  bool use_pdl = true;
  bool use_chain = true;
  bool use_2cta = true;   // encoder / prefill GEMMs through the cta_group::2 kernel (gemm2.cuh)
  // fp32-grade parity mode: every GEMM operand is a (hi, lo) bf16 pair and each GEMM computes a_hi w_hi + a_lo w_hi +
  // a_hi w_lo in ONE pass of the same tcgen05 kernel (activations stored [hi | lo | hi], weights [hi | hi | lo] along K);
  // attention, K/V caches and q/k/v stay fp32; exact QuickGELU.  ~3x the GEMM work: a verification mode (the north star's
  // "logits within 1e-3" against the fp32 reference), not a serving mode.  Weights must be (re-)uploaded after switching.
  bool tc_attn = true;    // ViT / prefill attention on tcgen05 (flash_attn_tc_kernel) when the sequence fits TMEM (S <= 512)
  bool use_mega = true;   // greedy decode steps of <= 64 sequences through the persistent decode_mega_kernel
  bool mega_coop = true;  // ... launched cooperatively (co-residency of its 148 CTAs guaranteed by the driver)
  bool mega_ready = false;
  int debug_layers = -1;  // debugging: run only the first n decoder layers in the decode step (both step paths); -1 = all
  bool parity = false;
  int ks() const { return parity ? 3 : 1; }                  // K multiplier of every GEMM operand
  size_t kvb() const { return parity ? 4 : 2; }              // bytes per K/V cache element
  const gitb200_engine* weights_from = nullptr;   // non-null: weight buffers are borrowed from that engine

  // derived geometry
  int g = 0, L = 0, Kpatch = 0, Kp = 0, d = 0, D = 0, F = 0, V = 0;
  // input size of the next encode (gitb200_set_input_size; default image_size x image_size): patch grid gh x gw,
  // Lc = gh * gw + 1 tokens per image; differs from (g, g, L) for MinMaxResizeForTest inputs (reference inference.py:29-64)
  int in_h = 0, in_w = 0, gh = 0, gw = 0, Lc = 0;

  // weights
  DevBuf w_patch, cls, pos_emb, lnpre_g, lnpre_b, lnpost_g, lnpost_b;
  std::vector<EncLayer> enc;
  DevBuf w_vp, b_vp, lnvp_g, lnvp_b, words_f32, words_bf16, positions, lnemb_g, lnemb_b, out_bias, temb;
  DevBuf m_lm;                                              // packed LM-head tiles (decode_mega_kernel)
  std::vector<DecLayer> dec;
  std::set<std::string> seen;
  bool finalized = false;

  // workspaces
  DevBuf x, h, qkv, ctx, u, feats, feats_f32, pos_interp;   // encoder
  DevBuf pt, pxd, phd, pq, pctx, pu;                        // prefill
  DevBuf img_kv, txt_kv, src_row[2];                        // caches
  DevBuf xd_t, hd_t, qkv_t, ctx_t, t_t, u_t, logits;        // decode step
  DevBuf y_t, qb_t, mega_bar;                               // decode_mega_kernel: pre-LN sums, bf16 q, grid-barrier counters
  DevBuf state, next_token, logprob_sum, tokens_i64, stage_img, stage_tok, stage_lp, prefix_dev;
  DevBuf beam_ws;                                           // beam-search bookkeeping (search.cuh)
  DevBuf sel_ws;                                            // greedy selection partials
  DevBuf chain;                                             // decode-step kernel chain completion counters [64]
  int attn_chunk_rows = 0, attn_box_rows = 0, attn_grid = 0;
  size_t attn_smem = 0;
  int cur_B = 0, cur_frames = 0, cur_M = 0, cur_beam = 1, T_alloc = 0, cur_rows = 0, cur_src = 0;

  EncodeTiledFn encode_tiled = nullptr;
  std::map<TmapKey, CUtensorMap> tmaps;

  // decode-step graph cache
  // (a call's graph is keyed by its row count, cache geometry and buffer addresses; several shapes alternate when batches
  //  are coalesced into launches of different sizes, so a handful of instantiated graphs are kept)
  struct StepGraph { cudaGraphExec_t exec = nullptr; int64_t launches = 0; unsigned long long last_use = 0; };
  std::map<std::vector<long long>, StepGraph> step_graphs;
  unsigned long long step_graph_clock = 0;
  int last_gemm_grid = 0;
  cudaStream_t own_stream = nullptr;
  cudaEvent_t own_event = nullptr;
  // asynchronous generate: enqueue now, read the loop state back in gitb200_generate_finish
  StepState* host_state = nullptr;   // pinned
  bool pending = false;
  int pend_max_steps = 0;
  bool pend_beam = false;
  cudaStream_t pend_stream = nullptr;
  cudaEvent_t chunk_ev[2] = {nullptr, nullptr};   // decode-loop chunks (generate_impl)
  cudaEvent_t dec_ev[2] = {nullptr, nullptr};     // around the decode loop of the last generate (gitb200_last_decode_ms)
  int dec_steps = 0;                               // step launches between them
  bool dec_mega = false;                           // ... each of which was one decode_mega_kernel launch
  // per-row prefixes of the NEXT generate call (gitb200_set_row_prefixes; consumed by that call)
  const int64_t* rp_tok = nullptr;
  const int32_t* rp_lens = nullptr;
  int rp_rows = 0, rp_stride = 0;
  // vocabulary trie (gitb200_set_trie; sticky) and the uniforms of the next sampled generate (gitb200_set_sampling)
  DevBuf trie_begin, trie_token, trie_child, trie_cursor;
  int trie_nodes = 0;
  const float* sample_u = nullptr;
  int sample_steps = 0, sample_rows = 0;
  float sample_temperature = 1.0f;
  bool constrained = false;            // this call's greedy selection runs constrained_select_kernel
  int64_t* pend_tok_host = nullptr;  // host-buffer variant: results land here
};

// Row range of one decode chain (the whole batch; kept as a struct so that step_layers reports its chain tail).
struct Lane {
  int row0 = 0, rows = 0;   // sequences (images * beam)
  int b0 = 0, nb = 0;       // images
  cudaStream_t st = nullptr;
  int chain_idx = 0;        // out: chain position / CTAs of the last kernel launched by step_layers
  unsigned int chain_ctas = 0;
};

static void drop_step_graphs(gitb200_engine* h) {
  for (auto& kv : h->step_graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  h->step_graphs.clear();
}

static int fail(gitb200_engine* h, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (h) h->err = buf; else g_create_error = buf;
  return 1;
}

#define CK(call)                                                                                     \
  do {                                                                                               \
    cudaError_t e_ = (call);                                                                         \
    if (e_ != cudaSuccess) return fail(h, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)
#define CKL(h_, what)                                                                                \
  do {                                                                                               \
    cudaError_t e_ = cudaGetLastError();                                                             \
    if (e_ != cudaSuccess) return fail(h_, "launch %s failed: %s", what, cudaGetErrorString(e_));    \
    (h_)->launches++;                                                                                \
  } while (0)
#define TRY(expr)                 \
  do {                            \
    int rc_ = (expr);             \
    if (rc_ != 0) return rc_;     \
  } while (0)

// Kernel launch with optional programmatic dependent launch (the decode step chains ~45 small kernels; PDL lets
// each one's prologue / weight prefetch overlap its predecessor's tail, also inside a captured CUDA graph).
template <typename... KArgs, typename... Args>
static cudaError_t launch_k(bool pdl, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                            Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

// ------------------------------------------------------------------------------------------------
// TMA descriptors
// ------------------------------------------------------------------------------------------------
static int load_encode_fn(gitb200_engine* h) {
  if (h->encode_tiled) return 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || fn == nullptr)
    return fail(h, "cuTensorMapEncodeTiled not available from the driver: %s", cudaGetErrorString(e));
  h->encode_tiled = reinterpret_cast<EncodeTiledFn>(fn);
  return 0;
}

// bf16 matrix [rows, cols] with leading dimension ld (elements); box = [box_rows x 64], 128B-swizzled (GEMM
// operands) or plain row-major (swizzle == false: decode-attention K/V slices).
static int get_tmap(gitb200_engine* h, const void* ptr, long long rows, long long cols, long long ld, int box_rows,
                    CUtensorMap* out, bool swizzle = true) {
  TmapKey key{ptr, rows, cols, ld, swizzle ? box_rows : -box_rows};
  auto it = h->tmaps.find(key);
  if (it != h->tmaps.end()) {
    *out = it->second;
    return 0;
  }
  TRY(load_encode_fn(h));
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || (ld * 2) % 16 != 0)
    return fail(h, "TMA operand must be 16-byte aligned with a 16-byte multiple row pitch (ptr=%p ld=%lld)", ptr, ld);
  CUtensorMap m;
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {64u, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = h->encode_tiled(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(h, "cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld cols=%lld ld=%lld box=%d)",
                                     static_cast<int>(r), rows, cols, ld, box_rows);
  if (h->tmaps.size() > 4096) h->tmaps.clear();
  h->tmaps[key] = m;
  *out = m;
  return 0;
}

// ------------------------------------------------------------------------------------------------
// GEMM launcher
// ------------------------------------------------------------------------------------------------
struct GemmCall {
  const bf16* A = nullptr;  // [M, K] (kernel operand A: 128-row tiles)
  long long lda = 0;
  const bf16* B = nullptr;  // [N, K] (kernel operand B: BN-row tiles)
  long long ldb = 0;
  GemmParams p{};
  int bn = 0;               // 0 = heuristic
};

template <int BN, int EPI>
static int launch_gemm_inst(gitb200_engine* h, const GemmCall& c, cudaStream_t st) {
  using C = GemmCfg<BN>;
  static bool attr_set[64] = {false};
  if (!attr_set[h->device & 63]) {
    CK(cudaFuncSetAttribute(gemm_bf16_tcgen05<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_set[h->device & 63] = true;
  }
  CUtensorMap ta, tb;
  TRY(get_tmap(h, c.A, c.p.M, c.p.K, c.lda, 128, &ta));
  TRY(get_tmap(h, c.B, c.p.N, c.p.K, c.ldb, BN, &tb));
  const int m_tiles = (c.p.M + 127) / 128;
  const int n_tiles = (c.p.N + BN - 1) / BN;
  const int tiles = m_tiles * n_tiles * c.p.k_splits;
  const int grid = tiles < h->num_sms ? tiles : h->num_sms;
  h->last_gemm_grid = grid;
  CK(launch_k(c.p.pdl != 0, gemm_bf16_tcgen05<BN, EPI>, dim3(grid), dim3(C::THREADS), C::SMEM_BYTES, st, ta, tb, c.p));
  CKL(h, "gemm_bf16_tcgen05");
  return 0;
}

template <int BN, int EPI>
static int launch_gemm2_inst(gitb200_engine* h, const GemmCall& c, cudaStream_t st) {
  using C = Gemm2Cfg<BN>;
  static bool attr_set[64] = {false};
  if (!attr_set[h->device & 63]) {
    CK(cudaFuncSetAttribute(gemm2_bf16_tcgen05<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_set[h->device & 63] = true;
  }
  CUtensorMap ta, tb;
  TRY(get_tmap(h, c.A, c.p.M, c.p.K, c.lda, 128, &ta));
  TRY(get_tmap(h, c.B, c.p.N, c.p.K, c.ldb, BN / 2, &tb));
  const int m_tiles = (c.p.M + 255) / 256;
  const int n_tiles = (c.p.N + BN - 1) / BN;
  const int tiles = m_tiles * n_tiles;
  const int pairs = std::min(tiles, h->num_sms / 2);
  h->last_gemm_grid = 2 * pairs;
  CK(launch_k(false, gemm2_bf16_tcgen05<BN, EPI>, dim3(2 * pairs), dim3(C::THREADS), C::SMEM_BYTES, st, ta, tb, c.p));
  CKL(h, "gemm2_bf16_tcgen05");
  return 0;
}
// The epilogue variants the hot path uses (each is its own kernel instantiation).
template <int BN>
static int launch_gemm2_bn(gitb200_engine* h, const GemmCall& c, cudaStream_t st) {
  const GemmParams& p = c.p;
  switch (epi_code(false, p.out_bf16 != 0, p.resid != nullptr, false, p.act)) {
    case epi_code(false, true, false, false, ACT_NONE): return launch_gemm2_inst<BN, epi_code(false, true, false, false, ACT_NONE)>(h, c, st);
    case epi_code(false, false, true, false, ACT_NONE): return launch_gemm2_inst<BN, epi_code(false, false, true, false, ACT_NONE)>(h, c, st);
    case epi_code(false, false, false, false, ACT_NONE): return launch_gemm2_inst<BN, epi_code(false, false, false, false, ACT_NONE)>(h, c, st);
    case epi_code(false, true, false, false, ACT_QUICKGELU): return launch_gemm2_inst<BN, epi_code(false, true, false, false, ACT_QUICKGELU)>(h, c, st);
    case epi_code(false, true, false, false, ACT_GELU_ERF): return launch_gemm2_inst<BN, epi_code(false, true, false, false, ACT_GELU_ERF)>(h, c, st);
    default: break;
  }
  return fail(h, "gemm2: epilogue combination not instantiated");
}

template <int BN>
static int launch_gemm_bn(gitb200_engine* h, const GemmCall& c, cudaStream_t st) {
  const GemmParams& p = c.p;
  const int code = epi_code(p.transposed != 0, p.out_bf16 != 0, p.resid != nullptr, p.partial != 0, p.act) | (p.split3 ? EPI_SPLIT3 : 0);
  if constexpr (BN == 256) {   // parity mode: bf16 outputs that feed another GEMM leave as [hi | lo | hi]
    switch (code) {
      case epi_code(false, true, false, false, ACT_QUICKGELU_EXACT) | EPI_SPLIT3: return launch_gemm_inst<BN, epi_code(false, true, false, false, ACT_QUICKGELU_EXACT) | EPI_SPLIT3>(h, c, st);
      case epi_code(false, true, false, false, ACT_GELU_ERF) | EPI_SPLIT3: return launch_gemm_inst<BN, epi_code(false, true, false, false, ACT_GELU_ERF) | EPI_SPLIT3>(h, c, st);
      default: break;
    }
  }
  if constexpr (BN == 64 || BN == 128 || BN == 256) {
    if (code == (epi_code(true, true, false, false, ACT_GELU_ERF) | EPI_SPLIT3))
      return launch_gemm_inst<BN, epi_code(true, true, false, false, ACT_GELU_ERF) | EPI_SPLIT3>(h, c, st);
  }
  if constexpr (BN == 192 || BN == 256 || BN == 128) {
    switch (code) {
      case epi_code(false, true, false, false, ACT_NONE): return launch_gemm_inst<BN, epi_code(false, true, false, false, ACT_NONE)>(h, c, st);
      case epi_code(false, false, true, false, ACT_NONE): return launch_gemm_inst<BN, epi_code(false, false, true, false, ACT_NONE)>(h, c, st);
      case epi_code(false, false, false, false, ACT_NONE): return launch_gemm_inst<BN, epi_code(false, false, false, false, ACT_NONE)>(h, c, st);
      case epi_code(false, true, false, false, ACT_QUICKGELU): return launch_gemm_inst<BN, epi_code(false, true, false, false, ACT_QUICKGELU)>(h, c, st);
      case epi_code(false, true, false, false, ACT_GELU_ERF): return launch_gemm_inst<BN, epi_code(false, true, false, false, ACT_GELU_ERF)>(h, c, st);
      default: break;
    }
  }
  if constexpr (BN == 64 || BN == 128 || BN == 256) {
    switch (code) {
      case epi_code(true, false, false, true, ACT_NONE): return launch_gemm_inst<BN, epi_code(true, false, false, true, ACT_NONE)>(h, c, st);
      case epi_code(true, false, false, false, ACT_NONE): return launch_gemm_inst<BN, epi_code(true, false, false, false, ACT_NONE)>(h, c, st);
      case epi_code(true, true, false, false, ACT_GELU_ERF): return launch_gemm_inst<BN, epi_code(true, true, false, false, ACT_GELU_ERF)>(h, c, st);
      default: break;
    }
  }
  return fail(h, "gemm: epilogue combination not instantiated (transposed=%d bf16=%d resid=%d partial=%d act=%d bn=%d)",
              p.transposed, p.out_bf16, p.resid != nullptr, p.partial, p.act, BN);
}

static int pick_bn(const gitb200_engine* h, int M, int N, bool transposed) {
  if (transposed) return N <= 64 ? 64 : (N <= 128 ? 128 : 256);
  // Measured on B200 (tools/gemm_sweep.py, M = 12608): wide outputs (N = 2304 / 3072) run best with 128x256 tiles
  // (fewest operand bytes per FLOP through L2/shared memory); N = 768 with 128x192 tiles (4 column tiles: less
  // wave quantisation than 3 x 256 at 99 row tiles over 148 SMs).
  (void)h; (void)M;
  if (N % 256 == 0 && N >= 1024) return 256;
  if (N % 192 == 0) return 192;
  if (N % 256 == 0) return 256;
  if (N % 128 == 0) return 128;
  return N > 192 ? 256 : (N > 128 ? 192 : 128);
}

static int launch_gemm(gitb200_engine* h, GemmCall c, cudaStream_t st) {
  GemmParams& p = c.p;
  if (p.k_splits < 1) p.k_splits = 1;
  const int kb_total = (p.K + 63) / 64;
  if (p.k_splits > kb_total) p.k_splits = kb_total;
  {
    const int kb_per = (kb_total + p.k_splits - 1) / p.k_splits;
    p.k_splits = (kb_total + kb_per - 1) / kb_per;  // no empty split
  }
  if (p.seg_n <= 0) p.seg_n = p.N;
  if (p.rows_per_batch <= 0) {
    p.rows_per_batch = p.M;
    p.batch_stride = p.M;
  }
  if (!p.transposed && (p.N % 32 != 0 || p.seg_n % 32 != 0))
    return fail(h, "gemm: N and segment width must be multiples of 32 (N=%d seg=%d)", p.N, p.seg_n);
  if (p.partial && !p.transposed) return fail(h, "gemm: split-K partial buffers are only implemented for the transposed epilogue");
  if (p.k_splits > 1 && !p.partial) return fail(h, "gemm: k_splits > 1 needs the partial-sum epilogue");
  if (p.partial && p.split_stride < static_cast<long long>(p.N) * p.ldo) return fail(h, "gemm: split_stride smaller than one partial buffer");
  int bn = c.bn > 0 ? c.bn : pick_bn(h, p.M, p.N, p.transposed != 0);
  if (p.split3 && !p.transposed) bn = 256;
  // 2-CTA (cta_group::2) kernel: explicit request (bn = 1000 + BN, unit tests) or engine option for the big GEMMs
  // Measured (tools/gemm_sweep.py, M = 12608): pairs win on wide outputs (+7-10 %) and on K = 3072 (+11 %); the
  // K = N = 768 out-projection is epilogue bound and stays on 1-CTA 128x192 tiles.
  if (bn < 1000 && h->use_2cta && !h->parity && !p.transposed && p.k_splits == 1 && p.M >= 2048) {
    if (p.N % 256 == 0 && p.N >= 1024) bn = 1256;
    else if (p.N % 192 == 0 && p.K >= 1536) bn = 1192;
  }
  if (bn >= 1000) {
    if (p.transposed || p.k_splits != 1) return fail(h, "gemm2: normal epilogue, no split-K");
    switch (bn - 1000) {
      case 256: return launch_gemm2_bn<256>(h, c, st);
      case 192: return launch_gemm2_bn<192>(h, c, st);
      default: return fail(h, "gemm2: unsupported tile width %d", bn - 1000);
    }
  }
  switch (bn) {
    case 64: return launch_gemm_bn<64>(h, c, st);
    case 128: return launch_gemm_bn<128>(h, c, st);
    case 192: return launch_gemm_bn<192>(h, c, st);
    case 256: return launch_gemm_bn<256>(h, c, st);
    default: return fail(h, "gemm: unsupported tile width %d", bn);
  }
}

// Plain C = A W^T (+bias)(+act)(+resid) -> out (fp32 or bf16), identity row map.
static GemmCall gemm_plain(const bf16* A, long long lda, const bf16* W, long long ldw, int M, int N, int K,
                           const float* bias, int act, const float* resid, void* out, bool out_bf16) {
  GemmCall c;
  c.A = A; c.lda = lda; c.B = W; c.ldb = ldw;
  c.p.M = M; c.p.N = N; c.p.K = K; c.p.k_splits = 1;
  c.p.bias = bias; c.p.act = act; c.p.resid = resid; c.p.ld_resid = N;
  c.p.out[0] = out; c.p.ldo = N; c.p.out_bf16 = out_bf16 ? 1 : 0;
  c.p.seg_n = N;
  return c;
}
// Skinny decode-step GEMM: out[r][f] = sum_k X[r][k] W[f][k] (+bias[f]) (+act) -- swap-AB, transposed epilogue.
// k_splits > 1: split s writes its partial sums to out + s * rows * ldo (fp32); the consumer adds them in split order.
static GemmCall gemm_skinny(const bf16* X, long long ldx, const bf16* W, long long ldw, int rows, int feats, int K,
                            const float* bias, int act, void* out, long long ldo, bool out_bf16, int k_splits,
                            const int* skip, bool pdl = false) {
  GemmCall c;
  c.A = W; c.lda = ldw; c.B = X; c.ldb = ldx;
  c.p.M = feats; c.p.N = rows; c.p.K = K; c.p.k_splits = k_splits;
  c.p.transposed = 1; c.p.partial = k_splits > 1 ? 1 : 0;
  c.p.split_stride = static_cast<long long>(rows) * ldo;
  c.p.bias = bias; c.p.act = act;
  c.p.out[0] = out; c.p.ldo = ldo; c.p.out_bf16 = out_bf16 ? 1 : 0;
  c.p.skip = skip;
  c.p.pdl = pdl ? 1 : 0;
  return c;
}

// ------------------------------------------------------------------------------------------------
// other launch helpers
// ------------------------------------------------------------------------------------------------
static int launch_ln(gitb200_engine* h, const LnParams& p, int D, cudaStream_t st, bool pdl = false) {
  const int grid = (p.rows + 7) / 8;
  if (D == 768 && pdl) CK(launch_k(true, layernorm_kernel<768, true>, dim3(grid), dim3(256), 0, st, p));
  else if (D == 768) CK(launch_k(false, layernorm_kernel<768, false>, dim3(grid), dim3(256), 0, st, p));
  else if (D == 1024) CK(launch_k(pdl, layernorm_kernel<1024, false>, dim3(grid), dim3(256), 0, st, p));
  else return fail(h, "layernorm: unsupported width %d", D);
  CKL(h, "layernorm_kernel");
  return 0;
}
static LnParams ln_params(const float* x, const float* bias, const float* resid, const float* g, const float* b,
                          float eps, float* of32, bf16* obf16, int rows) {
  LnParams p{};
  p.x = x; p.bias = bias; p.resid = resid; p.gamma = g; p.beta = b; p.eps = eps;
  p.out_f32 = of32; p.out_bf16 = obf16; p.rows = rows;
  return p;
}

template <int NW>
static void launch_flash(const AttnParams& p, cudaStream_t st) {
  dim3 grid((p.S + NW * 16 - 1) / (NW * 16), p.H, p.B);
  flash_attn_kernel<NW><<<grid, NW * 32, 0, st>>>(p);
}
// tcgen05 attention (attention.cuh: flash_attn_tc_kernel) for sequences whose scores fit the tensor memory in one piece
static int launch_attention_tc(gitb200_engine* h, const AttnParams& ap, cudaStream_t st) {
  AttnTcParams p{};
  p.out = ap.out; p.B = ap.B; p.S = ap.S; p.H = ap.H;
  p.spad = (ap.S + 15) / 16 * 16;
  if (p.spad <= 256) { p.kv_boxes = 1; p.kv_box_rows = p.spad; }
  else { p.kv_boxes = 2; p.kv_box_rows = ((p.spad + 1) / 2 + 7) / 8 * 8; }
  p.q_rows_per_batch = ap.S; p.kv_rows_per_batch = ap.S;
  p.q_col0 = 0; p.k_col0 = 0; p.v_col0 = 0;
  p.o_rs = ap.o_rs; p.o_bs = ap.o_bs;
  p.scale_log2 = 0.125f * 1.44269504088896340736f;
  const long long rows = static_cast<long long>(ap.B) * ap.S;
  CUtensorMap tq, tk, tv;
  TRY(get_tmap(h, ap.q, rows, ap.H * 64, ap.q_rs, 128, &tq));
  TRY(get_tmap(h, ap.k, rows, ap.H * 64, ap.kv_rs, p.kv_box_rows, &tk));
  TRY(get_tmap(h, ap.v, rows, ap.H * 64, ap.kv_rs, p.kv_box_rows, &tv));
  const size_t smem = attn_tc_smem_bytes(p.spad, p.kv_box_rows, p.kv_boxes);
  static size_t attr_done[64] = {0};
  if (attr_done[h->device & 63] < smem) {
    CK(cudaFuncSetAttribute(flash_attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    attr_done[h->device & 63] = smem;
  }
  flash_attn_tc_kernel<<<dim3(ap.H, ap.B), kAttnTcThreads, smem, st>>>(tq, tk, tv, p);
  CKL(h, "flash_attn_tc_kernel");
  return 0;
}

// tcgen05 attention for sequences beyond the tensor memory (attention.cuh: flash_attn_tc_long_kernel): 128-key blocks, two passes
static int launch_attention_tc_long(gitb200_engine* h, const AttnParams& ap, cudaStream_t st) {
  AttnTcParams p{};
  p.out = ap.out; p.B = ap.B; p.S = ap.S; p.H = ap.H;
  p.spad = (ap.S + 15) / 16 * 16;
  p.kv_boxes = 1; p.kv_box_rows = kAttnLongBlk;
  p.q_rows_per_batch = ap.S; p.kv_rows_per_batch = ap.S;
  p.q_col0 = 0; p.k_col0 = 0; p.v_col0 = 0;
  p.o_rs = ap.o_rs; p.o_bs = ap.o_bs;
  p.scale_log2 = 0.125f * 1.44269504088896340736f;
  const long long rows = static_cast<long long>(ap.B) * ap.S;
  CUtensorMap tq, tk, tv;
  TRY(get_tmap(h, ap.q, rows, ap.H * 64, ap.q_rs, 128, &tq));
  TRY(get_tmap(h, ap.k, rows, ap.H * 64, ap.kv_rs, kAttnLongBlk, &tk));
  TRY(get_tmap(h, ap.v, rows, ap.H * 64, ap.kv_rs, kAttnLongBlk, &tv));
  const size_t smem = attn_tc_long_smem_bytes();
  static bool attr_done[64] = {false};
  if (!attr_done[h->device & 63]) {
    CK(cudaFuncSetAttribute(flash_attn_tc_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    attr_done[h->device & 63] = true;
  }
  flash_attn_tc_long_kernel<<<dim3(ap.H, ap.B, (ap.S + 127) / 128), kAttnTcThreads, smem, st>>>(tq, tk, tv, p);
  CKL(h, "flash_attn_tc_long_kernel");
  return 0;
}

static int launch_attention(gitb200_engine* h, const AttnParams& ap, cudaStream_t st) {
  if (h->tc_attn && ap.S <= 512 && ap.q_bs == static_cast<long long>(ap.S) * ap.q_rs && ap.kv_bs == static_cast<long long>(ap.S) * ap.kv_rs &&
      attn_tc_smem_bytes((ap.S + 15) / 16 * 16, ap.S <= 256 ? (ap.S + 15) / 16 * 16 : (((ap.S + 15) / 16 * 16 + 1) / 2 + 7) / 8 * 8,
                         ap.S <= 256 ? 1 : 2) <= 226 * 1024)
    return launch_attention_tc(h, ap, st);
  if (h->tc_attn && ap.S > 512 && ap.q_bs == static_cast<long long>(ap.S) * ap.q_rs && ap.kv_bs == static_cast<long long>(ap.S) * ap.kv_rs)
    return launch_attention_tc_long(h, ap, st);
  AttnParams p = ap;
  p.scale_log2 = 0.125f * 1.44269504088896340736f;
  // query rows per CTA = 16 * NW: least padding first, then the larger tile (K/V are re-read per query tile)
  const int cands[4] = {8, 7, 6, 4};
  int best = 4;
  long long best_pad = 1LL << 60;
  for (int i = 0; i < 4; ++i) {
    const int rows = cands[i] * 16;
    const long long padded = static_cast<long long>((p.S + rows - 1) / rows) * rows;
    if (padded < best_pad) { best_pad = padded; best = cands[i]; }
  }
  switch (best) {
    case 8: launch_flash<8>(p, st); break;
    case 7: launch_flash<7>(p, st); break;
    case 6: launch_flash<6>(p, st); break;
    default: launch_flash<4>(p, st); break;
  }
  CKL(h, "flash_attn_kernel");
  return 0;
}

static int launch_attention_f32(gitb200_engine* h, const AttnF32Params& p, cudaStream_t st) {
  const size_t smem = static_cast<size_t>(4) * (64 + p.S) * sizeof(float);
  if (smem > 200 * 1024) return fail(h, "parity attention: %d keys do not fit in shared memory", p.S);
  static size_t attr_done[64] = {0};
  if (attr_done[h->device & 63] < smem) {
    CK(cudaFuncSetAttribute(attn_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    attr_done[h->device & 63] = smem;
  }
  const long long items = static_cast<long long>(p.B) * p.H * p.S;
  attn_f32_kernel<<<static_cast<unsigned int>((items + 3) / 4), 128, smem, st>>>(p);
  CKL(h, "attn_f32_kernel");
  return 0;
}

__global__ void cvt_rows_kernel(const float* __restrict__ src, long long src_ld, bf16* __restrict__ dst, long long dst_ld,
                                long long rows, long long cols, long long dst_cols) {
  const long long total = rows * dst_cols;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / dst_cols, c = i - r * dst_cols;
    dst[r * dst_ld + c] = __float2bfloat16_rn(c < cols ? src[r * src_ld + c] : 0.0f);
  }
}
// parity mode weights: [rows, 3 * dst_cols] = [hi | hi | lo]
__global__ void cvt_rows_split3_kernel(const float* __restrict__ src, long long src_ld, bf16* __restrict__ dst, long long rows,
                                       long long cols, long long dst_cols) {
  const long long total = rows * dst_cols;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / dst_cols, c = i - r * dst_cols;
    bf16 hi, lo;
    split_bf16(c < cols ? src[r * src_ld + c] : 0.0f, hi, lo);
    bf16* d = dst + r * 3 * dst_cols + c;
    d[0] = hi; d[dst_cols] = hi; d[2 * dst_cols] = lo;
  }
}
__global__ void sum_partials_kernel(const float* __restrict__ parts, float* __restrict__ out, long long n, int splits) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float a = parts[i];
    for (int s = 1; s < splits; ++s) a += parts[s * n + i];
    out[i] = a;
  }
}
__global__ void set_state_kernel(StepState* st, int pos, int cur_len, unsigned int* chain) {
  st->pos = pos; st->cur_len = cur_len; st->finished = 0; st->final_len = cur_len; st->step = 0;
  st->empty_caption = 0; st->ticket = 0; st->not_eos = 0; st->error = 0;
  for (int k = 0; k < 64; ++k) chain[k] = 0;
}
__global__ void init_generate_kernel(long long* tokens_out, long long* next_token, float* logprob_sum,
                                     const long long* prefix, int P, int rows, int max_steps, int sos, long long prefix_row_stride) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const long long* pr = prefix ? prefix + r * prefix_row_stride : nullptr;   // stride 0: one prefix for all rows
  for (int i = 0; i < P; ++i) tokens_out[static_cast<long long>(r) * max_steps + i] = pr ? pr[i] : sos;
  next_token[r] = pr ? pr[0] : sos;
  logprob_sum[r] = 0.f;
}
__global__ void advance_prefix_kernel(StepState* st, long long* next_token, const long long* prefix, int idx, int rows,
                                      unsigned int* chain) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < rows) next_token[r] = prefix[idx];
  if (r == 0) st->pos = st->pos + 1;
  if (r < 64) chain[r] = 0;
}

// ------------------------------------------------------------------------------------------------
// create / destroy / weights
// ------------------------------------------------------------------------------------------------
extern "C" int gitb200_abi_version(void) { return GITB200_ABI_VERSION; }

extern "C" const char* gitb200_last_error(const gitb200_engine* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

extern "C" int64_t gitb200_launch_count(const gitb200_engine* h) { return h ? h->launches : 0; }

extern "C" int gitb200_set_option(gitb200_engine* h, const char* name, int64_t value) {
  if (!h || !name) return 1;
  // the captured decode-step graph bakes the launch configuration in: drop it whenever an option changes
  drop_step_graphs(h);
  if (strcmp(name, "use_graph") == 0) { h->use_graph = value != 0; return 0; }
  if (strcmp(name, "use_pdl") == 0) { h->use_pdl = value != 0; return 0; }
  if (strcmp(name, "use_chain") == 0) { h->use_chain = value != 0; return 0; }
  if (strcmp(name, "use_2cta") == 0) { h->use_2cta = value != 0; return 0; }
  if (strcmp(name, "use_mega") == 0) { h->use_mega = value != 0; return 0; }
  if (strcmp(name, "tc_attn") == 0) { h->tc_attn = value != 0; return 0; }
  if (strcmp(name, "debug_layers") == 0) { h->debug_layers = static_cast<int>(value); return 0; }
  if (strcmp(name, "mega_coop") == 0) { h->mega_coop = value != 0; return 0; }
  if (strcmp(name, "parity") == 0) {
    if (h->weights_from != nullptr) return fail(h, "parity: this engine borrows its weights; switch the owning engine");
    if (h->parity != (value != 0)) { h->parity = value != 0; h->finalized = false; h->seen.clear(); h->tmaps.clear(); }
    return 0;
  }
  return fail(h, "unknown option %s", name);
}

extern "C" int gitb200_set_input_size(gitb200_engine* h, int height, int width) {
  if (!h) return 1;
  if (h->pending) return fail(h, "set_input_size: a generate call is in flight");
  const int p = h->cfg.patch;
  if (height < p || width < p) return fail(h, "set_input_size: %dx%d is smaller than one %dx%d patch", height, width, p, p);
  const long long tokens = static_cast<long long>(height / p) * (width / p) + 1;
  if (tokens > 16384) return fail(h, "set_input_size: %dx%d gives %lld tokens per image (limit 16384)", height, width, tokens);
  h->in_h = height;
  h->in_w = width;
  h->gh = height / p;      // nn.Conv2d(kernel = stride = patch, no padding) drops a trailing partial patch
  h->gw = width / p;
  h->Lc = h->gh * h->gw + 1;
  return 0;
}

extern "C" int gitb200_create(const gitb200_config* cfg, int device, gitb200_engine** out) {
  gitb200_engine* h = nullptr;
  if (!cfg || !out) return fail(nullptr, "gitb200_create: null argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(nullptr, "gitb200_create: no CUDA device visible (this engine has no CPU path)");
  if (device < 0 || device >= ndev) return fail(nullptr, "gitb200_create: bad device %d", device);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return fail(nullptr, "cudaGetDeviceProperties failed");
  if (prop.major != 10)
    return fail(nullptr, "gitb200_create: device %d is sm_%d%d; this library contains sm_100a code only", device, prop.major, prop.minor);
  if (cfg->image_size % cfg->patch != 0) return fail(nullptr, "image_size %% patch != 0");
  if (cfg->enc_width != 768 && cfg->enc_width != 1024) return fail(nullptr, "enc_width must be 768 or 1024");
  if (cfg->dec_hidden != 768 || cfg->dec_heads * 64 != cfg->dec_hidden || cfg->enc_heads * 64 != cfg->enc_width)
    return fail(nullptr, "head dim must be 64 and dec_hidden 768");
  if (cfg->dec_ffn % 64 != 0) return fail(nullptr, "dec_ffn must be a multiple of 64");
  h = new gitb200_engine();
  h->cfg = *cfg;
  h->device = device;
  h->num_sms = prop.multiProcessorCount;
  h->g = cfg->image_size / cfg->patch;
  h->L = h->g * h->g + 1;
  h->in_h = h->in_w = cfg->image_size;
  h->gh = h->gw = h->g;
  h->Lc = h->L;
  h->Kpatch = 3 * cfg->patch * cfg->patch;
  h->Kp = (h->Kpatch + 63) / 64 * 64;
  h->d = cfg->enc_width;
  h->D = cfg->dec_hidden;
  h->F = cfg->dec_ffn;
  h->V = cfg->vocab;
  h->enc.resize(cfg->enc_layers);
  h->dec.resize(cfg->dec_layers);
  cudaSetDevice(device);
  *out = h;
  return 0;
}

static void release_all(gitb200_engine* h) {
  DevBuf* bufs[] = {&h->w_patch, &h->cls, &h->pos_emb, &h->lnpre_g, &h->lnpre_b, &h->lnpost_g, &h->lnpost_b, &h->w_vp,
                    &h->b_vp, &h->lnvp_g, &h->lnvp_b, &h->words_f32, &h->words_bf16, &h->positions, &h->lnemb_g,
                    &h->lnemb_b, &h->out_bias, &h->temb, &h->m_lm, &h->y_t, &h->qb_t, &h->mega_bar, &h->x, &h->h, &h->qkv, &h->ctx, &h->u, &h->feats, &h->feats_f32, &h->pos_interp,
                    &h->pt, &h->pxd, &h->phd, &h->pq, &h->pctx, &h->pu, &h->img_kv, &h->txt_kv, &h->src_row[0],
                    &h->src_row[1], &h->xd_t, &h->hd_t, &h->qkv_t, &h->ctx_t, &h->t_t, &h->u_t, &h->logits, &h->state,
                    &h->next_token, &h->logprob_sum, &h->tokens_i64, &h->stage_img, &h->stage_tok, &h->stage_lp,
                    &h->prefix_dev, &h->beam_ws, &h->sel_ws, &h->chain};
  for (DevBuf* b : bufs) b->release();
  for (auto& l : h->enc) {
    DevBuf* lb[] = {&l.wqkv, &l.bqkv, &l.wo, &l.bo, &l.ln1g, &l.ln1b, &l.ln2g, &l.ln2b, &l.w1, &l.b1, &l.w2, &l.b2};
    for (DevBuf* b : lb) b->release();
  }
  for (auto& l : h->dec) {
    DevBuf* lb[] = {&l.wqkv, &l.bqkv, &l.wo, &l.bo, &l.lnag, &l.lnab, &l.w1, &l.b1, &l.w2, &l.b2, &l.lnog, &l.lnob,
                    &l.m_wqkv, &l.m_wo, &l.m_w1, &l.m_w2};
    for (DevBuf* b : lb) b->release();
  }
}

// Weight buffers of an engine, in a fixed order (the same for every engine of one geometry).
static std::vector<DevBuf*> weight_bufs(gitb200_engine* h) {
  std::vector<DevBuf*> v = {&h->w_patch, &h->cls, &h->pos_emb, &h->lnpre_g, &h->lnpre_b, &h->lnpost_g, &h->lnpost_b, &h->w_vp,
                            &h->b_vp, &h->lnvp_g, &h->lnvp_b, &h->words_f32, &h->words_bf16, &h->positions, &h->lnemb_g,
                            &h->lnemb_b, &h->out_bias, &h->temb, &h->m_lm};
  for (auto& l : h->enc)
    for (DevBuf* b : {&l.wqkv, &l.bqkv, &l.wo, &l.bo, &l.ln1g, &l.ln1b, &l.ln2g, &l.ln2b, &l.w1, &l.b1, &l.w2, &l.b2}) v.push_back(b);
  for (auto& l : h->dec)
    for (DevBuf* b : {&l.wqkv, &l.bqkv, &l.wo, &l.bo, &l.lnag, &l.lnab, &l.w1, &l.b1, &l.w2, &l.b2, &l.lnog, &l.lnob,
                      &l.m_wqkv, &l.m_wo, &l.m_w1, &l.m_w2}) v.push_back(b);
  return v;
}

extern "C" int gitb200_share_weights(gitb200_engine* h, gitb200_engine* src) {
  if (!h || !src) return 1;
  if (h == src) return fail(h, "share_weights: an engine cannot borrow from itself");
  if (!src->finalized) return fail(h, "share_weights: the source engine's weights are not finalized");
  if (src->weights_from != nullptr) return fail(h, "share_weights: the source engine borrows its weights itself");
  if (h->device != src->device) return fail(h, "share_weights: engines live on different devices");
  if (memcmp(&h->cfg, &src->cfg, sizeof(gitb200_config)) != 0) return fail(h, "share_weights: geometries differ");
  if (h->pending) return fail(h, "share_weights: a generate call is in flight");
  std::vector<DevBuf*> dst = weight_bufs(h), from = weight_bufs(src);
  for (size_t i = 0; i < dst.size(); ++i) dst[i]->borrow(*from[i]);
  h->tmaps.clear();
  drop_step_graphs(h);
  h->seen = src->seen;
  h->parity = src->parity;
  h->mega_ready = src->mega_ready;
  h->finalized = true;
  h->weights_from = src;
  return 0;
}

extern "C" void gitb200_destroy(gitb200_engine* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  drop_step_graphs(h);
  if (h->host_state) cudaFreeHost(h->host_state);
  if (h->own_event) cudaEventDestroy(h->own_event);
  for (int i = 0; i < 2; ++i) if (h->chunk_ev[i]) cudaEventDestroy(h->chunk_ev[i]);
  for (int i = 0; i < 2; ++i) if (h->dec_ev[i]) cudaEventDestroy(h->dec_ev[i]);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  release_all(h);
  delete h;
}

// Copy an fp32 source tensor into an engine buffer: fp32 (as is) or bf16 [rows, dst_cols] with zero padding,
// at a row offset inside the destination (used to fuse q/k/v into one matrix).
static int store_f32(gitb200_engine* h, DevBuf& dst, size_t total_elems, size_t elem_off, const float* src, size_t n,
                     cudaStream_t st) {
  CK(dst.ensure(total_elems * sizeof(float)));
  CK(cudaMemcpyAsync(dst.as<float>() + elem_off, src, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return 0;
}
static int store_bf16(gitb200_engine* h, DevBuf& dst, long long total_rows, long long dst_cols, long long row_off,
                      const float* src, long long rows, long long cols, cudaStream_t st) {
  CK(dst.ensure(static_cast<size_t>(total_rows) * dst_cols * h->ks() * sizeof(bf16)));
  const long long total = rows * dst_cols;
  const int grid = static_cast<int>(std::min<long long>((total + 255) / 256, 148 * 16));
  if (h->parity) cvt_rows_split3_kernel<<<grid, 256, 0, st>>>(src, cols, dst.as<bf16>() + row_off * 3 * dst_cols, rows, cols, dst_cols);
  else cvt_rows_kernel<<<grid, 256, 0, st>>>(src, cols, dst.as<bf16>() + row_off * dst_cols, dst_cols, rows, cols, dst_cols);
  CKL(h, "cvt_rows_kernel");
  return 0;
}

static bool shape_is(const int64_t* s, int nd, std::initializer_list<int64_t> want) {
  if (nd != static_cast<int>(want.size())) return false;
  int i = 0;
  for (int64_t w : want) if (s[i++] != w) return false;
  return true;
}

extern "C" int gitb200_set_weight(gitb200_engine* h, const char* ref_key, const void* dev_ptr, const int64_t* shape,
                                  int ndim, int dtype, void* stream) {
  if (!h) return 1;
  if (!ref_key || !dev_ptr || !shape) return fail(h, "set_weight: null argument");
  if (dtype != GITB200_F32) return fail(h, "set_weight(%s): only fp32 sources are accepted", ref_key);
  cudaSetDevice(h->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const float* src = static_cast<const float*>(dev_ptr);
  const std::string key(ref_key);
  const int d = h->d, D = h->D, F = h->F, V = h->V, L = h->L;
  auto bad_shape = [&]() { return fail(h, "set_weight(%s): unexpected shape", ref_key); };
  if (h->weights_from != nullptr) return fail(h, "set_weight(%s): this engine borrows its weights (gitb200_share_weights)", ref_key);
  h->finalized = false;
  int layer = -1;
  char sub[128];
  if (key == "image_encoder.proj" || key == "textual.output.weight") { h->seen.insert(key); return 0; }
  if (key == "image_encoder.class_embedding") {
    if (!shape_is(shape, ndim, {d})) return bad_shape();
    TRY(store_f32(h, h->cls, d, 0, src, d, st));
  } else if (key == "image_encoder.positional_embedding") {
    if (!shape_is(shape, ndim, {L, d})) return bad_shape();
    TRY(store_f32(h, h->pos_emb, static_cast<size_t>(L) * d, 0, src, static_cast<size_t>(L) * d, st));
  } else if (key == "image_encoder.conv1.weight") {
    if (!shape_is(shape, ndim, {d, 3, h->cfg.patch, h->cfg.patch})) return bad_shape();
    TRY(store_bf16(h, h->w_patch, d, h->Kp, 0, src, d, h->Kpatch, st));
  } else if (key == "image_encoder.ln_pre.weight") { if (!shape_is(shape, ndim, {d})) return bad_shape(); TRY(store_f32(h, h->lnpre_g, d, 0, src, d, st));
  } else if (key == "image_encoder.ln_pre.bias") { if (!shape_is(shape, ndim, {d})) return bad_shape(); TRY(store_f32(h, h->lnpre_b, d, 0, src, d, st));
  } else if (key == "image_encoder.ln_post.weight") { if (!shape_is(shape, ndim, {d})) return bad_shape(); TRY(store_f32(h, h->lnpost_g, d, 0, src, d, st));
  } else if (key == "image_encoder.ln_post.bias") { if (!shape_is(shape, ndim, {d})) return bad_shape(); TRY(store_f32(h, h->lnpost_b, d, 0, src, d, st));
  } else if (sscanf(ref_key, "image_encoder.transformer.resblocks.%d.%127s", &layer, sub) == 2) {
    if (layer < 0 || layer >= static_cast<int>(h->enc.size())) return fail(h, "set_weight(%s): layer out of range", ref_key);
    EncLayer& l = h->enc[layer];
    const std::string s(sub);
    if (s == "attn.in_proj_weight") { if (!shape_is(shape, ndim, {3 * d, d})) return bad_shape(); TRY(store_bf16(h, l.wqkv, 3 * d, d, 0, src, 3 * d, d, st)); }
    else if (s == "attn.in_proj_bias") { if (!shape_is(shape, ndim, {3 * d})) return bad_shape(); TRY(store_f32(h, l.bqkv, 3 * d, 0, src, 3 * d, st)); }
    else if (s == "attn.out_proj.weight") { if (!shape_is(shape, ndim, {d, d})) return bad_shape(); TRY(store_bf16(h, l.wo, d, d, 0, src, d, d, st)); }
    else if (s == "attn.out_proj.bias") { if (!shape_is(shape, ndim, {d})) return bad_shape(); TRY(store_f32(h, l.bo, d, 0, src, d, st)); }
    else if (s == "ln_1.weight") { if (!shape_is(shape, ndim, {d})) return bad_shape(); TRY(store_f32(h, l.ln1g, d, 0, src, d, st)); }
    else if (s == "ln_1.bias") { if (!shape_is(shape, ndim, {d})) return bad_shape(); TRY(store_f32(h, l.ln1b, d, 0, src, d, st)); }
    else if (s == "ln_2.weight") { if (!shape_is(shape, ndim, {d})) return bad_shape(); TRY(store_f32(h, l.ln2g, d, 0, src, d, st)); }
    else if (s == "ln_2.bias") { if (!shape_is(shape, ndim, {d})) return bad_shape(); TRY(store_f32(h, l.ln2b, d, 0, src, d, st)); }
    else if (s == "mlp.c_fc.weight") { if (!shape_is(shape, ndim, {4 * d, d})) return bad_shape(); TRY(store_bf16(h, l.w1, 4 * d, d, 0, src, 4 * d, d, st)); }
    else if (s == "mlp.c_fc.bias") { if (!shape_is(shape, ndim, {4 * d})) return bad_shape(); TRY(store_f32(h, l.b1, 4 * d, 0, src, 4 * d, st)); }
    else if (s == "mlp.c_proj.weight") { if (!shape_is(shape, ndim, {d, 4 * d})) return bad_shape(); TRY(store_bf16(h, l.w2, d, 4 * d, 0, src, d, 4 * d, st)); }
    else if (s == "mlp.c_proj.bias") { if (!shape_is(shape, ndim, {d})) return bad_shape(); TRY(store_f32(h, l.b2, d, 0, src, d, st)); }
    else return fail(h, "set_weight: unknown key %s", ref_key);
  } else if (key == "textual.visual_projection.0.weight") {
    if (!shape_is(shape, ndim, {D, d})) return bad_shape();
    TRY(store_bf16(h, h->w_vp, D, d, 0, src, D, d, st));
  } else if (key == "textual.visual_projection.0.bias") { if (!shape_is(shape, ndim, {D})) return bad_shape(); TRY(store_f32(h, h->b_vp, D, 0, src, D, st));
  } else if (key == "textual.visual_projection.1.weight") { if (!shape_is(shape, ndim, {D})) return bad_shape(); TRY(store_f32(h, h->lnvp_g, D, 0, src, D, st));
  } else if (key == "textual.visual_projection.1.bias") { if (!shape_is(shape, ndim, {D})) return bad_shape(); TRY(store_f32(h, h->lnvp_b, D, 0, src, D, st));
  } else if (key == "textual.embedding.words.weight") {
    if (!shape_is(shape, ndim, {V, D})) return bad_shape();
    TRY(store_f32(h, h->words_f32, static_cast<size_t>(V) * D, 0, src, static_cast<size_t>(V) * D, st));
    TRY(store_bf16(h, h->words_bf16, V, D, 0, src, V, D, st));
  } else if (key == "textual.embedding.positions.weight") {
    if (!shape_is(shape, ndim, {h->cfg.max_positions, D})) return bad_shape();
    TRY(store_f32(h, h->positions, static_cast<size_t>(h->cfg.max_positions) * D, 0, src, static_cast<size_t>(h->cfg.max_positions) * D, st));
  } else if (key == "textual.embedding.layer_norm.weight") { if (!shape_is(shape, ndim, {D})) return bad_shape(); TRY(store_f32(h, h->lnemb_g, D, 0, src, D, st));
  } else if (key == "textual.embedding.layer_norm.bias") { if (!shape_is(shape, ndim, {D})) return bad_shape(); TRY(store_f32(h, h->lnemb_b, D, 0, src, D, st));
  } else if (key == "textual.output.bias") { if (!shape_is(shape, ndim, {V})) return bad_shape(); TRY(store_f32(h, h->out_bias, V, 0, src, V, st));
  } else if (sscanf(ref_key, "textual.transformer.encoder.layer.%d.%127s", &layer, sub) == 2) {
    if (layer < 0 || layer >= static_cast<int>(h->dec.size())) return fail(h, "set_weight(%s): layer out of range", ref_key);
    DecLayer& l = h->dec[layer];
    const std::string s(sub);
    const char* qkvn[3] = {"query", "key", "value"};
    bool done = false;
    for (int i = 0; i < 3 && !done; ++i) {
      if (s == std::string("attention.self.") + qkvn[i] + ".weight") {
        if (!shape_is(shape, ndim, {D, D})) return bad_shape();
        TRY(store_bf16(h, l.wqkv, 3 * D, D, static_cast<long long>(i) * D, src, D, D, st));
        done = true;
      } else if (s == std::string("attention.self.") + qkvn[i] + ".bias") {
        if (!shape_is(shape, ndim, {D})) return bad_shape();
        TRY(store_f32(h, l.bqkv, 3 * D, static_cast<size_t>(i) * D, src, D, st));
        done = true;
      }
    }
    if (done) { /* stored */ }
    else if (s == "attention.output.dense.weight") { if (!shape_is(shape, ndim, {D, D})) return bad_shape(); TRY(store_bf16(h, l.wo, D, D, 0, src, D, D, st)); }
    else if (s == "attention.output.dense.bias") { if (!shape_is(shape, ndim, {D})) return bad_shape(); TRY(store_f32(h, l.bo, D, 0, src, D, st)); }
    else if (s == "attention.output.LayerNorm.weight") { if (!shape_is(shape, ndim, {D})) return bad_shape(); TRY(store_f32(h, l.lnag, D, 0, src, D, st)); }
    else if (s == "attention.output.LayerNorm.bias") { if (!shape_is(shape, ndim, {D})) return bad_shape(); TRY(store_f32(h, l.lnab, D, 0, src, D, st)); }
    else if (s == "intermediate.dense.weight") { if (!shape_is(shape, ndim, {F, D})) return bad_shape(); TRY(store_bf16(h, l.w1, F, D, 0, src, F, D, st)); }
    else if (s == "intermediate.dense.bias") { if (!shape_is(shape, ndim, {F})) return bad_shape(); TRY(store_f32(h, l.b1, F, 0, src, F, st)); }
    else if (s == "output.dense.weight") { if (!shape_is(shape, ndim, {D, F})) return bad_shape(); TRY(store_bf16(h, l.w2, D, F, 0, src, D, F, st)); }
    else if (s == "output.dense.bias") { if (!shape_is(shape, ndim, {D})) return bad_shape(); TRY(store_f32(h, l.b2, D, 0, src, D, st)); }
    else if (s == "output.LayerNorm.weight") { if (!shape_is(shape, ndim, {D})) return bad_shape(); TRY(store_f32(h, l.lnog, D, 0, src, D, st)); }
    else if (s == "output.LayerNorm.bias") { if (!shape_is(shape, ndim, {D})) return bad_shape(); TRY(store_f32(h, l.lnob, D, 0, src, D, st)); }
    else return fail(h, "set_weight: unknown key %s", ref_key);
  } else if (sscanf(ref_key, "img_temperal_embedding.%d", &layer) == 1) {
    if (layer < 0 || layer >= h->cfg.num_frames_emb) return fail(h, "set_weight(%s): frame out of range", ref_key);
    if (!shape_is(shape, ndim, {1, 1, d})) return bad_shape();
    TRY(store_f32(h, h->temb, static_cast<size_t>(h->cfg.num_frames_emb) * d, static_cast<size_t>(layer) * d, src, d, st));
  } else {
    return fail(h, "set_weight: unknown key %s", ref_key);
  }
  h->seen.insert(key);
  return 0;
}

extern "C" int gitb200_finalize_weights(gitb200_engine* h, void* stream) {
  if (!h) return 1;
  cudaSetDevice(h->device);
  std::vector<std::string> need = {"image_encoder.class_embedding", "image_encoder.positional_embedding",
                                   "image_encoder.conv1.weight", "image_encoder.ln_pre.weight", "image_encoder.ln_pre.bias",
                                   "image_encoder.ln_post.weight", "image_encoder.ln_post.bias",
                                   "textual.visual_projection.0.weight", "textual.visual_projection.0.bias",
                                   "textual.visual_projection.1.weight", "textual.visual_projection.1.bias",
                                   "textual.embedding.words.weight", "textual.embedding.positions.weight",
                                   "textual.embedding.layer_norm.weight", "textual.embedding.layer_norm.bias",
                                   "textual.output.bias"};
  const char* encs[] = {"attn.in_proj_weight", "attn.in_proj_bias", "attn.out_proj.weight", "attn.out_proj.bias",
                        "ln_1.weight", "ln_1.bias", "ln_2.weight", "ln_2.bias", "mlp.c_fc.weight", "mlp.c_fc.bias",
                        "mlp.c_proj.weight", "mlp.c_proj.bias"};
  for (int i = 0; i < h->cfg.enc_layers; ++i)
    for (const char* s : encs) need.push_back("image_encoder.transformer.resblocks." + std::to_string(i) + "." + s);
  const char* decs[] = {"attention.self.query.weight", "attention.self.query.bias", "attention.self.key.weight",
                        "attention.self.key.bias", "attention.self.value.weight", "attention.self.value.bias",
                        "attention.output.dense.weight", "attention.output.dense.bias", "attention.output.LayerNorm.weight",
                        "attention.output.LayerNorm.bias", "intermediate.dense.weight", "intermediate.dense.bias",
                        "output.dense.weight", "output.dense.bias", "output.LayerNorm.weight", "output.LayerNorm.bias"};
  for (int j = 0; j < h->cfg.dec_layers; ++j)
    for (const char* s : decs) need.push_back("textual.transformer.encoder.layer." + std::to_string(j) + "." + s);
  for (int f = 0; f < h->cfg.num_frames_emb; ++f) need.push_back("img_temperal_embedding." + std::to_string(f));
  for (const std::string& k : need)
    if (!h->seen.count(k)) return fail(h, "finalize_weights: missing tensor %s", k.c_str());
  // fragment-packed weight tiles of the persistent decode-step kernel (decode_mega.cuh)
  h->mega_ready = false;
  if (!h->parity && h->D == kMegaD && h->F == kMegaF && h->cfg.dec_heads == kMegaH && h->cfg.dec_layers <= 6) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    auto pack = [&](DevBuf& dst, const DevBuf& src, long long ldw, int n_feat, int k0, long long n_tiles, int stride, int offset) -> int {
      CK(dst.ensure(static_cast<size_t>(n_tiles) * stride * kMegaTileBytes));
      const long long total = n_tiles * 48 * 32;
      pack_tiles_kernel<<<static_cast<int>(std::min<long long>((total + 255) / 256, 148 * 16)), 256, 0, st>>>(
          src.as<bf16>(), ldw, n_feat, k0, dst.as<uint8_t>(), n_tiles, stride, offset);
      CKL(h, "pack_tiles_kernel");
      return 0;
    };
    for (auto& l : h->dec) {
      TRY(pack(l.m_wqkv, l.wqkv, h->D, 3 * h->D, 0, 3 * h->D / 8, 1, 0));
      TRY(pack(l.m_wo, l.wo, h->D, h->D, 0, h->D / 8, 1, 0));
      TRY(pack(l.m_w1, l.w1, h->D, h->F, 0, h->F / 8, 1, 0));
      for (int sl = 0; sl < 4; ++sl) TRY(pack(l.m_w2, l.w2, h->F, h->D, sl * kMegaD, h->D / 8, 4, sl));
    }
    TRY(pack(h->m_lm, h->words_bf16, h->D, h->V, 0, (h->V + 7) / 8, 1, 0));
    h->mega_ready = true;
  }
  CK(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
  CK(cudaGetLastError());
  h->finalized = true;
  return 0;
}

// ------------------------------------------------------------------------------------------------
// hot path A: encoder
// ------------------------------------------------------------------------------------------------
static int encode_impl(gitb200_engine* h, const float* images, int B, int frames, bool list_input, float* feats_out,
                       cudaStream_t st) {
  if (!h->finalized) return fail(h, "weights not finalized");
  if (B < 1 || frames < 1) return fail(h, "encode: bad batch/frames");
  if (list_input && h->cfg.num_frames_emb > 0 && frames > h->cfg.num_frames_emb) {
    // reference zip() truncates to the number of temporal embeddings (layers/decoder.py:848-849)
    frames = h->cfg.num_frames_emb;
  }
  const int d = h->d, L = h->Lc, gh = h->gh, gw = h->gw, Kp = h->Kp, H = h->cfg.enc_heads;
  const int NI = B * frames;
  const long long Me = static_cast<long long>(NI) * L;
  const int ks = h->ks();                 // parity mode: GEMM operands are [hi | lo | hi] -> 3x the K extent
  const bool par = h->parity;
  CK(h->x.ensure(Me * d * 4));
  CK(h->h.ensure(Me * d * 2 * ks));
  CK(h->qkv.ensure(Me * 3 * d * (par ? 4 : 2)));         // parity: q | k | v stay fp32
  CK(h->ctx.ensure(Me * d * 2 * ks));
  CK(h->u.ensure(std::max<long long>(Me * 4 * d * 2, static_cast<long long>(NI) * gh * gw * Kp * 2) * ks));
  CK(h->feats.ensure(Me * d * 2 * ks));
  float* x = h->x.as<float>();
  bf16* hb = h->h.as<bf16>();
  bf16* qkv = h->qkv.as<bf16>();
  bf16* ctx = h->ctx.as<bf16>();
  bf16* u = h->u.as<bf16>();

  // positional embedding of this input size: the stored one, or its bicubic re-sampling to the gh x gw grid
  // (reference layers/CLIP/model.py:245-251; recomputed per call: 1 + gh*gw rows, the parameters may have changed)
  const float* pos = h->pos_emb.as<float>();
  if (gh != h->g || gw != h->g) {
    CK(h->pos_interp.ensure(static_cast<size_t>(L) * d * 4));
    const long long total = static_cast<long long>(L) * (d / 4);
    const int grid = static_cast<int>(std::min<long long>((total + 255) / 256, h->num_sms * 8));
    pos_embed_bicubic_kernel<<<grid, 256, 0, st>>>(h->pos_emb.as<float>(), h->pos_interp.as<float>(), h->g, gh, gw, d);
    CKL(h, "pos_embed_bicubic_kernel");
    pos = h->pos_interp.as<float>();
  }
  // patch embedding: im2col + GEMM, rows land at token index 1 + patch (CLS row is filled by the next kernel)
  {
    const long long total = static_cast<long long>(NI) * gh * gw * (Kp / 8);
    const int grid = static_cast<int>(std::min<long long>((total + 255) / 256, h->num_sms * 16));
    im2col_patch_kernel<<<grid, 256, 0, st>>>(images, u, NI, h->in_h, h->in_w, h->cfg.patch, gh, gw, Kp, par ? 1 : 0);
    CKL(h, "im2col_patch_kernel");
    GemmCall c = gemm_plain(u, Kp * ks, h->w_patch.as<bf16>(), Kp * ks, NI * gh * gw, d, Kp * ks, nullptr, ACT_NONE, nullptr, x, false);
    c.p.rows_per_batch = gh * gw;
    c.p.batch_stride = L;
    c.p.row_offset = 1;
    TRY(launch_gemm(h, c, st));
    const int gridr = static_cast<int>((Me + 7) / 8);
    if (d == 768)
      cls_pos_lnpre_kernel<768><<<gridr, 256, 0, st>>>(x, h->cls.as<float>(), pos, h->lnpre_g.as<float>(), h->lnpre_b.as<float>(), static_cast<int>(Me), L);
    else
      cls_pos_lnpre_kernel<1024><<<gridr, 256, 0, st>>>(x, h->cls.as<float>(), pos, h->lnpre_g.as<float>(), h->lnpre_b.as<float>(), static_cast<int>(Me), L);
    CKL(h, "cls_pos_lnpre_kernel");
  }
  auto ln_enc = [&](const float* g, const float* b) {
    LnParams p = ln_params(x, nullptr, nullptr, g, b, 1e-5f, nullptr, hb, static_cast<int>(Me));
    p.split3 = par ? 1 : 0;
    return p;
  };
  for (int i = 0; i < h->cfg.enc_layers; ++i) {
    EncLayer& l = h->enc[i];
    TRY(launch_ln(h, ln_enc(l.ln1g.as<float>(), l.ln1b.as<float>()), d, st));
    TRY(launch_gemm(h, gemm_plain(hb, d * ks, l.wqkv.as<bf16>(), d * ks, static_cast<int>(Me), 3 * d, d * ks, l.bqkv.as<float>(), ACT_NONE, nullptr, qkv, !par), st));
    if (par) {
      AttnF32Params ap{};
      const float* qf = h->qkv.as<float>();
      ap.q = qf; ap.k = qf + d; ap.v = qf + 2 * d; ap.out = ctx;
      ap.B = NI; ap.S = L; ap.H = H; ap.d_model = d;
      ap.q_rs = 3 * d; ap.kv_rs = 3 * d; ap.q_bs = static_cast<long long>(L) * 3 * d; ap.kv_bs = ap.q_bs;
      ap.o_bs = static_cast<long long>(L) * 3 * d;
      TRY(launch_attention_f32(h, ap, st));
    } else {
      AttnParams ap{};
      ap.q = qkv; ap.k = qkv + d; ap.v = qkv + 2 * d; ap.out = ctx;
      ap.B = NI; ap.S = L; ap.H = H;
      ap.q_rs = 3 * d; ap.kv_rs = 3 * d; ap.q_bs = static_cast<long long>(L) * 3 * d; ap.kv_bs = ap.q_bs;
      ap.o_rs = d; ap.o_bs = static_cast<long long>(L) * d;
      TRY(launch_attention(h, ap, st));
    }
    TRY(launch_gemm(h, gemm_plain(ctx, d * ks, l.wo.as<bf16>(), d * ks, static_cast<int>(Me), d, d * ks, l.bo.as<float>(), ACT_NONE, x, x, false), st));
    TRY(launch_ln(h, ln_enc(l.ln2g.as<float>(), l.ln2b.as<float>()), d, st));
    {
      GemmCall c = gemm_plain(hb, d * ks, l.w1.as<bf16>(), d * ks, static_cast<int>(Me), 4 * d, d * ks, l.b1.as<float>(),
                              par ? ACT_QUICKGELU_EXACT : ACT_QUICKGELU, nullptr, u, true);
      if (par) { c.p.split3 = 1; c.p.ldo = 3LL * 4 * d; }
      TRY(launch_gemm(h, c, st));
    }
    TRY(launch_gemm(h, gemm_plain(u, 4 * d * ks, l.w2.as<bf16>(), 4 * d * ks, static_cast<int>(Me), d, 4 * d * ks, l.b2.as<float>(), ACT_NONE, x, x, false), st));
  }
  // ln_post on all tokens (+ temporal embedding), re-ordered to [B, frames*L, d]
  {
    LnParams p = ln_params(x, nullptr, nullptr, h->lnpost_g.as<float>(), h->lnpost_b.as<float>(), 1e-5f, feats_out, h->feats.as<bf16>(), static_cast<int>(Me));
    p.remap_B = B; p.remap_F = frames; p.remap_L = L;
    // temporal embeddings only for list inputs (reference layers/decoder.py:846-849; a bare tensor skips them)
    p.temb = (list_input && h->cfg.num_frames_emb > 0) ? h->temb.as<float>() : nullptr;
    p.split3 = par ? 1 : 0;
    TRY(launch_ln(h, p, d, st));
  }
  h->cur_B = B;
  h->cur_frames = frames;
  h->cur_M = frames * L;
  return 0;
}

// ------------------------------------------------------------------------------------------------
// hot path B: prefill (image rows of the decoder, computed once) + decode step
// ------------------------------------------------------------------------------------------------
// Split-K factors of the decode-step GEMMs: a handful of activation rows against [features, K] weights is latency
// bound, so K is spread over enough CTAs that each one has all of its weight tiles in flight at once.  Every split
// stores its own partial-sum buffer; the consumer (decode attention / LayerNorm) adds them in split order.
constexpr int kQkvSplits = 3, kOutProjSplits = 6, kFc2Splits = 8, kMaxProjSplits = 8;

// K/V caches: bf16, or fp32 in parity mode (void*: the element size is the engine's kvb())
static char* img_kv_ptr(gitb200_engine* h, int layer, int kv, long long elem_off = 0) {
  const long long per = static_cast<long long>(h->cur_B) * h->cur_M * h->D;
  return h->img_kv.as<char>() + ((static_cast<long long>(layer) * 2 + kv) * per + elem_off) * h->kvb();
}
static char* txt_kv_ptr(gitb200_engine* h, int layer, int kv, long long elem_off = 0) {
  const long long per = static_cast<long long>(h->cur_rows) * h->T_alloc * h->D;
  return h->txt_kv.as<char>() + ((static_cast<long long>(layer) * 2 + kv) * per + elem_off) * h->kvb();
}

static int prefill_impl(gitb200_engine* h, int B, int beam, int T_alloc, float* vproj_out, cudaStream_t st) {
  if (B != h->cur_B || h->cur_M <= 0) return fail(h, "prefill: call encode with the same batch first");
  const int D = h->D, F = h->F, d = h->d, M = h->cur_M, nl = h->cfg.dec_layers, H = h->cfg.dec_heads;
  const long long rows = static_cast<long long>(B) * M;
  const int R = B * beam;
  const int ks = h->ks();
  const bool par = h->parity;
  const long long kvb = static_cast<long long>(h->kvb());
  CK(h->pt.ensure(rows * D * 4));
  CK(h->pxd.ensure(rows * D * 4));
  CK(h->phd.ensure(rows * D * 2 * ks));
  CK(h->pq.ensure(rows * D * kvb));
  CK(h->pctx.ensure(rows * D * 2 * ks));
  CK(h->pu.ensure(rows * F * 2 * ks));
  CK(h->img_kv.ensure(static_cast<long long>(nl) * 2 * rows * D * kvb));
  {
    // decode_mega_kernel fetches whole 64-position boxes of the text cache and masks the positions past the caption's end
    // by giving them probability 0 -- which only works if what lies there is finite: a fresh allocation is zeroed once
    // (afterwards the buffer only ever holds K/V values or zeros)
    const void* before = h->txt_kv.p;
    CK(h->txt_kv.ensure(static_cast<long long>(nl) * 2 * R * T_alloc * D * kvb));
    if (h->txt_kv.p != before) CK(cudaMemsetAsync(h->txt_kv.p, 0, h->txt_kv.cap, st));
  }
  CK(h->src_row[0].ensure(static_cast<size_t>(R) * T_alloc * 4));
  CK(h->src_row[1].ensure(static_cast<size_t>(R) * T_alloc * 4));
  CK(h->xd_t.ensure(static_cast<size_t>(R) * D * 4));
  CK(h->hd_t.ensure(static_cast<size_t>(R) * D * 2 * ks));
  CK(h->qkv_t.ensure(static_cast<size_t>(kQkvSplits) * R * 3 * D * 4));   // split-K partial-sum buffers
  CK(h->ctx_t.ensure(static_cast<size_t>(R) * D * 2 * ks));
  CK(h->t_t.ensure(static_cast<size_t>(kMaxProjSplits) * R * D * 4));
  CK(h->u_t.ensure(static_cast<size_t>(R) * F * 2 * ks));
  CK(h->logits.ensure(static_cast<size_t>(R) * h->V * 4));
  CK(h->state.ensure(sizeof(StepState) + 64));            // + the megakernel's error word
  CK(h->y_t.ensure(static_cast<size_t>(R) * D * 4));
  CK(h->qb_t.ensure(static_cast<size_t>(R) * D * 2));
  CK(h->mega_bar.ensure(64));
  CK(h->next_token.ensure(static_cast<size_t>(R) * 8));
  CK(h->logprob_sum.ensure(static_cast<size_t>(R) * 4));
  h->cur_beam = beam;
  h->cur_rows = R;
  h->T_alloc = T_alloc;
  float* t = h->pt.as<float>();
  float* xd = h->pxd.as<float>();
  bf16* hd = h->phd.as<bf16>();
  bf16* q = h->pq.as<bf16>();
  bf16* ctx = h->pctx.as<bf16>();
  bf16* u = h->pu.as<bf16>();

  auto ln_pre = [&](const float* g, const float* b, float eps) {
    LnParams p = ln_params(t, nullptr, nullptr, g, b, eps, xd, hd, static_cast<int>(rows));
    p.split3 = par ? 1 : 0;
    return p;
  };
  // visual projection: Linear(dv -> 768) + LayerNorm(1e-5)
  TRY(launch_gemm(h, gemm_plain(h->feats.as<bf16>(), d * ks, h->w_vp.as<bf16>(), d * ks, static_cast<int>(rows), D, d * ks, h->b_vp.as<float>(), ACT_NONE, nullptr, t, false), st));
  TRY(launch_ln(h, ln_pre(h->lnvp_g.as<float>(), h->lnvp_b.as<float>(), 1e-5f), D, st));
  if (vproj_out) CK(cudaMemcpyAsync(vproj_out, xd, rows * D * 4, cudaMemcpyDeviceToDevice, st));
  for (int j = 0; j < nl; ++j) {
    DecLayer& l = h->dec[j];
    // fused q|k|v projection; k and v rows go straight into the image K/V cache (bf16; fp32 in parity mode)
    GemmCall c = gemm_plain(hd, D * ks, l.wqkv.as<bf16>(), D * ks, static_cast<int>(rows), 3 * D, D * ks, l.bqkv.as<float>(), ACT_NONE, nullptr, q, !par);
    c.p.seg_n = D;
    c.p.out[0] = q; c.p.out[1] = img_kv_ptr(h, j, 0); c.p.out[2] = img_kv_ptr(h, j, 1);
    c.p.ldo = D;
    TRY(launch_gemm(h, c, st));
    if (j + 1 == nl) break;  // image rows of the last layer are never read (text rows only need their K/V)
    if (par) {
      AttnF32Params ap{};
      ap.q = h->pq.as<float>(); ap.k = reinterpret_cast<const float*>(img_kv_ptr(h, j, 0));
      ap.v = reinterpret_cast<const float*>(img_kv_ptr(h, j, 1)); ap.out = ctx;
      ap.B = B; ap.S = M; ap.H = H; ap.d_model = D;
      ap.q_rs = D; ap.kv_rs = D; ap.q_bs = static_cast<long long>(M) * D; ap.kv_bs = ap.q_bs; ap.o_bs = 3 * ap.q_bs;
      TRY(launch_attention_f32(h, ap, st));
    } else {
      AttnParams ap{};
      ap.q = q; ap.k = reinterpret_cast<const bf16*>(img_kv_ptr(h, j, 0)); ap.v = reinterpret_cast<const bf16*>(img_kv_ptr(h, j, 1)); ap.out = ctx;
      ap.B = B; ap.S = M; ap.H = H;
      ap.q_rs = D; ap.kv_rs = D; ap.q_bs = static_cast<long long>(M) * D; ap.kv_bs = ap.q_bs; ap.o_rs = D; ap.o_bs = ap.q_bs;
      TRY(launch_attention(h, ap, st));
    }
    TRY(launch_gemm(h, gemm_plain(ctx, D * ks, l.wo.as<bf16>(), D * ks, static_cast<int>(rows), D, D * ks, l.bo.as<float>(), ACT_NONE, xd, t, false), st));
    TRY(launch_ln(h, ln_pre(l.lnag.as<float>(), l.lnab.as<float>(), 1e-12f), D, st));
    {
      GemmCall c1 = gemm_plain(hd, D * ks, l.w1.as<bf16>(), D * ks, static_cast<int>(rows), F, D * ks, l.b1.as<float>(), ACT_GELU_ERF, nullptr, u, true);
      if (par) { c1.p.split3 = 1; c1.p.ldo = 3LL * F; }
      TRY(launch_gemm(h, c1, st));
    }
    TRY(launch_gemm(h, gemm_plain(u, F * ks, l.w2.as<bf16>(), F * ks, static_cast<int>(rows), D, F * ks, l.b2.as<float>(), ACT_NONE, xd, t, false), st));
    TRY(launch_ln(h, ln_pre(l.lnog.as<float>(), l.lnob.as<float>(), 1e-12f), D, st));
  }
  return 0;
}

// One decode step for the `rows` sequences: embed next_token at state->pos, 6 layers against the KV caches,
// optional LM head -> h->logits.  Every kernel reads the position / finished flag from device state so the
// same launch sequence (and CUDA graph) serves every step.
static int step_layers(gitb200_engine* h, Lane& ln_, const long long* tokens, const int* src_row, bool lm_head) {
  const int D = h->D, F = h->F, R = ln_.rows, beam = h->cur_beam;
  const int nl = (h->debug_layers >= 0) ? std::min(h->debug_layers, h->cfg.dec_layers) : h->cfg.dec_layers;
  cudaStream_t st = ln_.st;
  StepState* state = h->state.as<StepState>();
  const int* skip = &state->finished;
  const int ks = h->ks();
  const bool par = h->parity;
  float* xd = h->xd_t.as<float>();
  bf16* hd = h->hd_t.as<bf16>();
  float* qkv = h->qkv_t.as<float>();
  bf16* ctx = h->ctx_t.as<bf16>();
  float* t = h->t_t.as<float>();
  bf16* u = h->u_t.as<bf16>();
  float* logits = h->logits.as<float>();
  const bool pdl = h->use_pdl;
  // Ordering inside the step: flag chain (greedy; the beam bookkeeping kernels still use grid dependencies).
  const bool chain_on = pdl && h->use_chain && beam == 1;
  ChainSync cs{};
  cs.counters = chain_on ? h->chain.as<unsigned int>() : nullptr;
  cs.idx = 0;
  cs.pred_ctas = 0;
  auto next_link = [&](unsigned int ctas_of_this_kernel) {  // call after each launch
    cs.idx += 1;
    cs.pred_ctas = ctas_of_this_kernel;
  };
  // chain head: launched WITHOUT the PDL attribute -> fully ordered after the previous step
  CK(launch_k(false, embed_ln_kernel<768>, dim3((R + 7) / 8), dim3(256), 0, st, tokens, 1LL, h->words_f32.as<float>(),
              h->positions.as<float>(), h->lnemb_g.as<float>(), h->lnemb_b.as<float>(), xd, hd, R, 0,
              static_cast<const StepState*>(state), h->V, par ? 1 : 0, cs));
  CKL(h, "embed_ln_kernel");
  next_link((R + 7) / 8);
  // split-K GEMMs write partial-sum buffers; bias / residual / LayerNorm live in the consumer kernel
  auto skinny = [&](GemmCall c) -> int {
    c.p.chain = cs;
    TRY(launch_gemm(h, c, st));
    next_link(static_cast<unsigned int>(h->last_gemm_grid));
    return 0;
  };
  auto ln_partials = [&](const float* parts, int n, int rows, const float* bias, const float* resid, const float* g,
                         const float* b, float* of32, bf16* obf16) {
    LnParams p = ln_params(parts, bias, resid, g, b, 1e-12f, of32, obf16, rows);
    p.n_partials = n;
    p.partial_stride = static_cast<long long>(rows) * D;
    p.split3 = par ? 1 : 0;
    return p;
  };
  auto ln = [&](LnParams p) -> int {
    p.skip_flag = skip; p.chain = cs;
    TRY(launch_ln(h, p, D, st, pdl));
    next_link((p.rows + 7) / 8);
    return 0;
  };
  for (int j = 0; j < nl; ++j) {
    DecLayer& l = h->dec[j];
    TRY(skinny(gemm_skinny(hd, D * ks, l.wqkv.as<bf16>(), D * ks, R, 3 * D, D * ks, nullptr, ACT_NONE, qkv, 3 * D, false, kQkvSplits, skip, pdl)));
    if (par) {
      DecAttnF32Params ap{};
      ap.qkv = qkv; ap.n_partials = kQkvSplits; ap.partial_stride = static_cast<long long>(R) * 3 * D;
      ap.bqkv = l.bqkv.as<float>();
      ap.img_k = reinterpret_cast<const float*>(img_kv_ptr(h, j, 0)); ap.img_v = reinterpret_cast<const float*>(img_kv_ptr(h, j, 1));
      ap.txt_k = reinterpret_cast<float*>(txt_kv_ptr(h, j, 0)); ap.txt_v = reinterpret_cast<float*>(txt_kv_ptr(h, j, 1));
      ap.src_row = src_row; ap.ctx = ctx; ap.R = R; ap.beam = beam; ap.M = h->cur_M; ap.T_alloc = h->T_alloc; ap.D = D;
      ap.state = state;
      ap.chain = cs;
      const size_t smem = static_cast<size_t>(4) * (192 + h->cur_M + h->T_alloc) * sizeof(float);
      if (smem > 200 * 1024) return fail(h, "parity decode attention: %d keys do not fit in shared memory", h->cur_M + h->T_alloc);
      CK(cudaFuncSetAttribute(decode_attn_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
      const unsigned int grid = static_cast<unsigned int>((R * h->cfg.dec_heads + 3) / 4);
      CK(launch_k(pdl, decode_attn_f32_kernel, dim3(grid), dim3(128), smem, st, ap));
      CKL(h, "decode_attn_f32_kernel");
      next_link(grid);
    } else {
      DecAttnParams ap{};
      ap.qkv = qkv; ap.n_partials = kQkvSplits; ap.partial_stride = static_cast<long long>(R) * 3 * D;
      ap.bqkv = l.bqkv.as<float>();
      ap.img_k = reinterpret_cast<const bf16*>(img_kv_ptr(h, j, 0)); ap.img_v = reinterpret_cast<const bf16*>(img_kv_ptr(h, j, 1));
      ap.txt_k = reinterpret_cast<bf16*>(txt_kv_ptr(h, j, 0)); ap.txt_v = reinterpret_cast<bf16*>(txt_kv_ptr(h, j, 1));
      ap.src_row = src_row; ap.ctx = ctx; ap.B = ln_.nb; ap.M = h->cur_M; ap.T_alloc = h->T_alloc; ap.D = D;
      ap.state = state;
      ap.chunk_rows = h->attn_chunk_rows; ap.box_rows = h->attn_box_rows;
      ap.chain = cs;
      CUtensorMap tk, tv;
      TRY(get_tmap(h, ap.img_k, static_cast<long long>(ln_.nb) * h->cur_M, D, D, ap.box_rows, &tk, false));
      TRY(get_tmap(h, ap.img_v, static_cast<long long>(ln_.nb) * h->cur_M, D, D, ap.box_rows, &tv, false));
      dim3 grid(std::min(h->attn_grid, ln_.nb * h->cfg.dec_heads));
      if (beam == 1) CK(launch_k(pdl, decode_attn_kernel<1, true>, grid, dim3(128), h->attn_smem, st, tk, tv, ap));
      else if (beam == 4) CK(launch_k(pdl, decode_attn_kernel<4>, grid, dim3(128), h->attn_smem, st, tk, tv, ap));
      else if (beam == 3) CK(launch_k(pdl, decode_attn_kernel<3>, grid, dim3(128), h->attn_smem, st, tk, tv, ap));
      else if (beam == 2) CK(launch_k(pdl, decode_attn_kernel<2>, grid, dim3(128), h->attn_smem, st, tk, tv, ap));
      else return fail(h, "decode: beam size %d not supported (1 .. 4)", beam);
      CKL(h, "decode_attn_kernel");
      next_link(grid.x);
    }
    TRY(skinny(gemm_skinny(ctx, D * ks, l.wo.as<bf16>(), D * ks, R, D, D * ks, nullptr, ACT_NONE, t, D, false, kOutProjSplits, skip, pdl)));
    TRY(ln(ln_partials(t, kOutProjSplits, R, l.bo.as<float>(), xd, l.lnag.as<float>(), l.lnab.as<float>(), xd, hd)));
    {
      GemmCall c1 = gemm_skinny(hd, D * ks, l.w1.as<bf16>(), D * ks, R, F, D * ks, l.b1.as<float>(), ACT_GELU_ERF, u, static_cast<long long>(F) * ks, true, 1, skip, pdl);
      c1.p.split3 = par ? 1 : 0;
      TRY(skinny(c1));
    }
    TRY(skinny(gemm_skinny(u, F * ks, l.w2.as<bf16>(), F * ks, R, D, F * ks, nullptr, ACT_NONE, t, D, false, kFc2Splits, skip, pdl)));
    TRY(ln(ln_partials(t, kFc2Splits, R, l.b2.as<float>(), xd, l.lnog.as<float>(), l.lnob.as<float>(), xd, hd)));
  }
  if (lm_head)
    TRY(skinny(gemm_skinny(hd, D * ks, h->words_bf16.as<bf16>(), D * ks, R, h->V, D * ks, h->out_bias.as<float>(), ACT_NONE, logits, h->V, false, 1, skip, pdl)));
  ln_.chain_idx = cs.idx;
  ln_.chain_ctas = cs.pred_ctas;
  return 0;
}

// Geometry of the decode-attention staging buffer: the image K/V slice of one (image, head) is M rows of 128 B;
// up to 512 rows are staged per round as 1-2 TMA boxes of <= 256 rows.
static int set_attn_smem_limit(gitb200_engine* h) {
  const int M = h->cur_M;
  // chunks of at most 224 keys: two (K + V) staging buffers of one chunk per CTA, and at least two CTAs per SM (M = 257
  // in one piece was 131 KB per CTA = one 4-warp CTA per SM: 44 us per launch at 128 beam rows, profiles/launches_r02_config3.csv)
  const int n_chunks = (M + 223) / 224;
  h->attn_box_rows = (M + n_chunks - 1) / n_chunks;
  h->attn_chunk_rows = h->attn_box_rows;
  h->attn_smem = static_cast<size_t>(4) * h->attn_chunk_rows * 128 + 128;
  int per_sm = static_cast<int>((227 * 1024) / (h->attn_smem + 8 * 1024));
  per_sm = std::max(1, std::min(per_sm, 4));
  const int items = h->cur_B * h->cfg.dec_heads;
  h->attn_grid = std::min(items, per_sm * h->num_sms);
  CK(cudaFuncSetAttribute(decode_attn_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(h->attn_smem)));
  CK(cudaFuncSetAttribute(decode_attn_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(h->attn_smem)));
  CK(cudaFuncSetAttribute(decode_attn_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(h->attn_smem)));
  CK(cudaFuncSetAttribute(decode_attn_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(h->attn_smem)));
  CK(h->chain.ensure(256));
  CK(cudaMemset(h->chain.p, 0, 256));
  return 0;
}

#include "engine_api.inc"
#include "preproc_api.inc"
