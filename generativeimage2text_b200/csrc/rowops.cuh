// Row-wise HBM-bound kernels of the GIT hot path: LayerNorm (+bias/+residual/+temporal embedding),
// patch im2col, CLS/pos-embed/ln_pre, token embedding + LN, and the greedy selection kernels.
// One warp per row, 128-bit loads/stores, fp32 statistics (two-pass: mean, then biased variance,
// like torch.nn.functional.layer_norm).
#pragma once
#include "ptx.cuh"

namespace gitb200 {

// Decode-loop state shared by the kernels of one generate() call (device memory, one per engine).
struct StepState {
  int pos;            // text position of the token fed at the current step (0-based, prefix included)
  int cur_len;        // tokens currently in tokens_out (start tokens + generated)
  int finished;       // greedy: every row ended with EOS -> remaining steps are no-ops
  int final_len;      // number of valid columns of tokens_out
  int step;           // number of decode steps executed so far
  int empty_caption;  // greedy step-0 special case (reference layers/decoder.py:279-291)
  unsigned int ticket;
  int not_eos;        // rows whose newest token is not EOS (per step, reset by the last block)
  int error;          // decode_mega_kernel: a bounded wait gave up (1 grid barrier, 2 ring consumer, 3 ring producer)
};

struct LnParams {
  const float* x;        // [rows, D] fp32 (ldx == D)
  const float* bias;     // [D] or null: added to x first
  const float* resid;    // [rows, D] or null: added to x first
  const float* gamma;
  const float* beta;
  float eps;
  float* out_f32;        // may alias x (in place) or be null
  __nv_bfloat16* out_bf16;  // or null
  int rows;
  int n_partials;        // > 1: x is the first of n split-K partial-sum buffers, partial_stride elements apart; they are
  long long partial_stride;  // added in split order (fixed order -> bit-reproducible; nothing to re-zero)
  // optional frame remap (video path, reference layers/decoder.py:846-851): input row (f*B + b)*L + l ->
  // output row b*(F*L) + f*L + l, plus `+ temb[f]` AFTER the normalisation.
  const float* temb;     // [F, D] or null
  int remap_B, remap_F, remap_L;  // remap_F == 0 -> identity
  const int* skip_flag;  // device int: non-zero -> kernel is a no-op (finished decode)
  int split3;            // parity mode: out_bf16 rows are [hi | lo | hi] (3 x D columns, see split_bf16 in ptx.cuh)
  ChainSync chain;       // decode-step flag ordering (counters == null: plain / PDL ordering)
};

// PRE: fetch gamma / beta / bias before the dependency wait (decode chain: hides one L2 round trip; costs registers,
// so the big encoder LayerNorms use PRE = false).
template <int D, bool PRE>
__global__ void __launch_bounds__(256) layernorm_kernel(const LnParams p) {
  static_assert(D % 128 == 0, "D");
  constexpr int NV = D / 128;
  griddep_launch_early();
  tl_mark(100002);
  const int lane = threadIdx.x & 31;
  // parameters are constants: fetch them before the dependency wait
  float4 gam[PRE ? NV : 1], bet[PRE ? NV : 1], bia[PRE ? NV : 1];
  if (PRE) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      gam[i] = __ldg(reinterpret_cast<const float4*>(p.gamma) + i * 32 + lane);
      bet[i] = __ldg(reinterpret_cast<const float4*>(p.beta) + i * 32 + lane);
      bia[i] = (p.bias != nullptr) ? __ldg(reinterpret_cast<const float4*>(p.bias) + i * 32 + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  const bool chained = p.chain.counters != nullptr;
  if (chained) {
    if (p.skip_flag != nullptr && *p.skip_flag != 0) return;  // stable within a step
    chain_wait(p.chain);
  } else {
    griddep_wait();
    if (p.skip_flag != nullptr && *p.skip_flag != 0) return;
  }
  tl_mark(2);
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= p.rows) {
    chain_signal(p.chain);
    return;
  }
  float4 v[NV];
  const float4* xp = reinterpret_cast<const float4*>(p.x + static_cast<long long>(row) * D);
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = __ldcg(xp + i * 32 + lane);
  for (int s0 = 1; s0 < p.n_partials; s0 += 3) {   // three partial buffers per round trip, added in split order
    float4 w[3][NV];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const bool ok = s0 + k < p.n_partials;
      const float4* sp = reinterpret_cast<const float4*>(p.x + (ok ? s0 + k : 0) * p.partial_stride + static_cast<long long>(row) * D);
#pragma unroll
      for (int i = 0; i < NV; ++i) w[k][i] = ok ? __ldcg(sp + i * 32 + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
      for (int i = 0; i < NV; ++i) { v[i].x += w[k][i].x; v[i].y += w[k][i].y; v[i].z += w[k][i].z; v[i].w += w[k][i].w; }
    }
  }
  if (p.resid != nullptr) {
    const float4* rp = reinterpret_cast<const float4*>(p.resid + static_cast<long long>(row) * D);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float4 r = __ldcg(rp + i * 32 + lane);
      v[i].x += r.x; v[i].y += r.y; v[i].z += r.z; v[i].w += r.w;
    }
  }
  if (PRE) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      v[i].x += bia[i].x; v[i].y += bia[i].y; v[i].z += bia[i].z; v[i].w += bia[i].w;
    }
  } else if (p.bias != nullptr) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias) + i * 32 + lane);
      v[i].x += b.x; v[i].y += b.y; v[i].z += b.z; v[i].w += b.w;
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  const float mean = warp_sum(s) * (1.0f / D);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    ss += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum(ss) * (1.0f / D) + p.eps);

  long long orow = row;
  int frame = 0;
  if (p.remap_F > 0) {
    const int img = row / p.remap_L;
    const int l = row - img * p.remap_L;
    frame = img / p.remap_B;
    const int b = img - frame * p.remap_B;
    orow = (static_cast<long long>(b) * p.remap_F + frame) * p.remap_L + l;
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float4 g = PRE ? gam[i] : __ldg(reinterpret_cast<const float4*>(p.gamma) + i * 32 + lane);
    const float4 b = PRE ? bet[i] : __ldg(reinterpret_cast<const float4*>(p.beta) + i * 32 + lane);
    float4 o;
    o.x = (v[i].x - mean) * rstd * g.x + b.x;
    o.y = (v[i].y - mean) * rstd * g.y + b.y;
    o.z = (v[i].z - mean) * rstd * g.z + b.z;
    o.w = (v[i].w - mean) * rstd * g.w + b.w;
    if (p.temb != nullptr) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(p.temb + static_cast<long long>(frame) * D) + i * 32 + lane);
      o.x += t.x; o.y += t.y; o.z += t.z; o.w += t.w;
    }
    if (p.out_f32 != nullptr) reinterpret_cast<float4*>(p.out_f32 + orow * D)[i * 32 + lane] = o;
    if (p.out_bf16 != nullptr && p.split3) {
      uint2 hi, lo;
      pack_split2(o.x, o.y, hi.x, lo.x);
      pack_split2(o.z, o.w, hi.y, lo.y);
      uint2* dst = reinterpret_cast<uint2*>(p.out_bf16 + orow * 3 * D) + i * 32 + lane;
      dst[0] = hi; dst[D / 4] = lo; dst[D / 2] = hi;
    } else if (p.out_bf16 != nullptr) {
      uint2 pk;
      pk.x = pack_bf16(o.x, o.y);
      pk.y = pack_bf16(o.z, o.w);
      reinterpret_cast<uint2*>(p.out_bf16 + orow * D)[i * 32 + lane] = pk;
    }
  }
  tl_mark(200002);
  chain_signal(p.chain);
}

// Patch im2col for the stride==kernel conv (reference layers/CLIP/model.py:224,242):
// A[(img, py, px)][(c, ky, kx)] = img[c, py*p+ky, px*p+kx], zero-padded to Kp columns, bf16.
__global__ void im2col_patch_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ A, int n_img, int H, int W,
                                    int p, int gh, int gw, int Kp, int split3) {
  const long long total = static_cast<long long>(n_img) * gh * gw * (Kp / 8);
  const int K = 3 * p * p;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int kc = static_cast<int>(idx % (Kp / 8));
    const long long row = idx / (Kp / 8);
    const int px = static_cast<int>(row % gw);
    const int py = static_cast<int>((row / gw) % gh);
    const long long im = row / (gh * gw);
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = kc * 8 + j;
      float x = 0.f;
      if (k < K) {
        const int c = k / (p * p);
        const int r = k - c * p * p;
        const int ky = r / p;
        const int kx = r - ky * p;
        x = __ldg(img + ((im * 3 + c) * H + (py * p + ky)) * static_cast<long long>(W) + (px * p + kx));
      }
      v[j] = x;
    }
    if (split3) {   // parity mode: [hi | lo | hi], 3 x Kp columns
      uint4 hi, lo;
      pack_split2(v[0], v[1], hi.x, lo.x);
      pack_split2(v[2], v[3], hi.y, lo.y);
      pack_split2(v[4], v[5], hi.z, lo.z);
      pack_split2(v[6], v[7], hi.w, lo.w);
      uint4* dst = reinterpret_cast<uint4*>(A + row * 3 * Kp) + kc;
      dst[0] = hi; dst[Kp / 8] = lo; dst[Kp / 4] = hi;
      continue;
    }
    uint4 o;
    o.x = pack_bf16(v[0], v[1]);
    o.y = pack_bf16(v[2], v[3]);
    o.z = pack_bf16(v[4], v[5]);
    o.w = pack_bf16(v[6], v[7]);
    reinterpret_cast<uint4*>(A + row * Kp)[kc] = o;
  }
}

// Positional embedding [1 + g0*g0, d] re-sampled to a gh x gw grid for inputs whose size differs from the resolution the
// embedding was built for (reference layers/CLIP/model.py:245-251): CLS row copied, grid rows =
// F.interpolate(mode='bicubic', align_corners=False), i.e. source coordinate (o + 0.5) * in / out - 0.5, Keys cubic with
// A = -0.75 over taps floor(x) - 1 .. floor(x) + 2 clamped to the grid. One thread per (token, 4 channels).
__device__ __forceinline__ void cubic_coeffs_a075(float t, float w[4]) {
  const float A = -0.75f;
  const float x0 = t + 1.0f, x1 = t, x2 = 1.0f - t, x3 = 2.0f - t;
  w[0] = ((A * x0 - 5.0f * A) * x0 + 8.0f * A) * x0 - 4.0f * A;
  w[1] = ((A + 2.0f) * x1 - (A + 3.0f)) * x1 * x1 + 1.0f;
  w[2] = ((A + 2.0f) * x2 - (A + 3.0f)) * x2 * x2 + 1.0f;
  w[3] = ((A * x3 - 5.0f * A) * x3 + 8.0f * A) * x3 - 4.0f * A;
}
__global__ void __launch_bounds__(256)
pos_embed_bicubic_kernel(const float* __restrict__ pos, float* __restrict__ out, int g0, int gh, int gw, int d) {
  const int d4 = d >> 2;
  const long long total = static_cast<long long>(1 + gh * gw) * d4;
  const float sy = static_cast<float>(g0) / static_cast<float>(gh), sx = static_cast<float>(g0) / static_cast<float>(gw);
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % d4);
    const int tok = static_cast<int>(idx / d4);
    const float4* src = reinterpret_cast<const float4*>(pos);
    float4 acc;
    if (tok == 0) {
      acc = __ldg(src + c);
    } else {
      const int oy = (tok - 1) / gw, ox = (tok - 1) - oy * gw;
      const float ry = sy * (oy + 0.5f) - 0.5f, rx = sx * (ox + 0.5f) - 0.5f;
      const float fy = floorf(ry), fx = floorf(rx);
      float wy[4], wx[4];
      cubic_coeffs_a075(ry - fy, wy);
      cubic_coeffs_a075(rx - fx, wx);
      const int iy = static_cast<int>(fy), ix = static_cast<int>(fx);
      acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int yy = min(max(iy - 1 + a, 0), g0 - 1);
        float4 row = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int xx = min(max(ix - 1 + b, 0), g0 - 1);
          const float4 v = __ldg(src + static_cast<long long>(1 + yy * g0 + xx) * d4 + c);
          row.x += wx[b] * v.x; row.y += wx[b] * v.y; row.z += wx[b] * v.z; row.w += wx[b] * v.w;
        }
        acc.x += wy[a] * row.x; acc.y += wy[a] * row.y; acc.z += wy[a] * row.z; acc.w += wy[a] * row.w;
      }
    }
    reinterpret_cast<float4*>(out)[idx] = acc;
  }
}

// x[img, l] = ln_pre((l == 0 ? class_embedding : patch_out[img, l]) + positional_embedding[l])
// in place on the fp32 residual stream (reference layers/CLIP/model.py:254-257).
template <int D>
__global__ void __launch_bounds__(256)
cls_pos_lnpre_kernel(float* __restrict__ x, const float* __restrict__ cls, const float* __restrict__ pos,
                     const float* __restrict__ gamma, const float* __restrict__ beta, int rows, int L) {
  constexpr int NV = D / 128;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int l = row % L;
  float4 v[NV];
  const float4* src = (l == 0) ? reinterpret_cast<const float4*>(cls)
                               : reinterpret_cast<const float4*>(x + static_cast<long long>(row) * D);
  const float4* pp = reinterpret_cast<const float4*>(pos + static_cast<long long>(l) * D);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float4 a = src[i * 32 + lane];
    const float4 b = __ldg(pp + i * 32 + lane);
    v[i] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mean = warp_sum(s) * (1.0f / D);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    ss += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum(ss) * (1.0f / D) + 1e-5f);
  float4* op = reinterpret_cast<float4*>(x + static_cast<long long>(row) * D);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + i * 32 + lane);
    const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + i * 32 + lane);
    op[i * 32 + lane] = make_float4((v[i].x - mean) * rstd * g.x + b.x, (v[i].y - mean) * rstd * g.y + b.y,
                                    (v[i].z - mean) * rstd * g.z + b.z, (v[i].w - mean) * rstd * g.w + b.w);
  }
}

// e = LN(words[tok] + positions[pos], eps 1e-8) (reference layers/decoder.py:65-78); one warp per row.
// tokens come from `tokens` (int64 [rows], stride tok_stride) ; position = pos_base + (state ? state->pos : 0).
template <int D>
__global__ void __launch_bounds__(256)
embed_ln_kernel(const long long* __restrict__ tokens, long long tok_stride, const float* __restrict__ words,
                const float* __restrict__ positions, const float* __restrict__ gamma, const float* __restrict__ beta,
                float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_bf16, int rows, int pos_base,
                const StepState* __restrict__ state, int vocab, int split3, const ChainSync chain) {
  constexpr int NV = D / 128;
  griddep_launch();
  griddep_wait();   // chain head: ordered after the previous step by a full dependency
  tl_mark(4);
  if (state != nullptr && state->finished) return;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) {
    chain_signal(chain);
    return;
  }
  const int lane = threadIdx.x & 31;
  long long tok = tokens[row * tok_stride];
  tok = tok < 0 ? 0 : (tok >= vocab ? vocab - 1 : tok);
  const int pos = pos_base + (state != nullptr ? state->pos : 0);
  const float4* wp = reinterpret_cast<const float4*>(words + tok * D);
  const float4* pp = reinterpret_cast<const float4*>(positions + static_cast<long long>(pos) * D);
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float4 a = __ldg(wp + i * 32 + lane);
    const float4 b = __ldg(pp + i * 32 + lane);
    v[i] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mean = warp_sum(s) * (1.0f / D);
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    ss += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum(ss) * (1.0f / D) + 1e-8f);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + i * 32 + lane);
    const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + i * 32 + lane);
    float4 o = make_float4((v[i].x - mean) * rstd * g.x + b.x, (v[i].y - mean) * rstd * g.y + b.y,
                           (v[i].z - mean) * rstd * g.z + b.z, (v[i].w - mean) * rstd * g.w + b.w);
    reinterpret_cast<float4*>(out_f32 + static_cast<long long>(row) * D)[i * 32 + lane] = o;
    if (split3) {
      uint2 hi, lo;
      pack_split2(o.x, o.y, hi.x, lo.x);
      pack_split2(o.z, o.w, hi.y, lo.y);
      uint2* dst = reinterpret_cast<uint2*>(out_bf16 + static_cast<long long>(row) * 3 * D) + i * 32 + lane;
      dst[0] = hi; dst[D / 4] = lo; dst[D / 2] = hi;
    } else {
      uint2 pk;
      pk.x = pack_bf16(o.x, o.y);
      pk.y = pack_bf16(o.z, o.w);
      reinterpret_cast<uint2*>(out_bf16 + static_cast<long long>(row) * D)[i * 32 + lane] = pk;
    }
  }
  chain_signal(chain);
}

// ------------------------------------------------------------------------------------------------
// Greedy selection = the body of AutoRegressiveBeamSearch.search for beam 1 / per-node 1
// (reference layers/decoder.py:258-273 first step, :313-417 loop): no-repeat scatter(-10000) on the
// input token (not on the first step), EOS forcing, log_softmax, argmax (lowest index on exact ties),
// logprob accumulation, all-EOS early exit.  One CTA per row.
// ------------------------------------------------------------------------------------------------
struct SelectParams {
  const float* logits;      // [rows, V]
  int V;
  int rows;
  int eos;
  int prefix_len;           // P
  int max_steps;
  long long* tokens_out;    // [rows, max_steps]
  float* logprob_sum;       // [rows]
  long long* next_token;    // [rows] input of the next step
  const long long* forced;  // [rows, max_steps] or null
  StepState* state;
  float* step_logits;       // optional dump [steps, rows_total, V]
  int rows_total, row0;     // this launch covers rows [row0, row0 + rows) of the batch
  // per-row prefixes (question batches, an extension of the reference's single-prefix path layers/decoder.py:985-1006):
  // row r starts from row_prefix[r * stride + 0 .. lens[r]); all rows advance in lockstep from text position 0, and while a
  // row is still inside its prefix its selection is overridden by the next prefix token (log-prob 0, no bookkeeping)
  const long long* row_prefix;
  int row_prefix_stride;
  const int* row_prefix_lens;
  // per-row partial results of the vocabulary slices (grid.x = n_split CTAs per row)
  int n_split;
  float* part_max;          // [rows, n_split]
  float* part_sum;          // [rows, n_split]  sum exp(v - part_max)
  int* part_arg;            // [rows, n_split]
  unsigned int* row_ticket; // [rows]
  ChainSync chain;          // last kernel of the chain: waits, then re-zeroes all counters when the step is over
};

// grid (n_split, rows): each CTA folds one vocabulary slice of one row into (max, argmax, sum exp) with a
// single online pass; the last CTA of a row combines the slices and does the reference's bookkeeping.
__global__ void __launch_bounds__(256) greedy_select_kernel(const SelectParams p) {
  griddep_launch_early();
  tl_mark(100005);
  StepState* st = p.state;
  if (p.chain.counters != nullptr) {
    if (st->finished) return;  // stable within a step
    chain_wait(p.chain);
  } else {
    griddep_wait();
    if (st->finished) return;
  }
  tl_mark(5);
  const int row = blockIdx.y;
  const int split = blockIdx.x;
  const int tid = threadIdx.x;
  const int step = st->step;
  const int cur_len = st->cur_len;
  const float* z = p.logits + static_cast<long long>(row) * p.V;
  // The input token of this step (== our previous choice unless teacher forcing is on): the reference's
  // masks are functions of the *input* sequence (predictions_so_far[:, -1]).
  const long long last = p.next_token[row];
  const int own_prefix = (p.row_prefix != nullptr) ? p.row_prefix_lens[p.row0 + row] : 0;
  const bool in_prefix = (p.row_prefix != nullptr) && cur_len < own_prefix;
  const bool first = (p.row_prefix != nullptr) ? (cur_len == own_prefix) : (step == 0);   // the row's first real decision
  const int chunk = (p.V + p.n_split - 1) / p.n_split;
  const int lo = split * chunk;
  const int hi = min(p.V, lo + chunk);
  if (p.step_logits != nullptr) {
    float* dst = p.step_logits + (static_cast<long long>(step) * p.rows_total + p.row0 + row) * p.V;
    for (int i = lo + tid; i < hi; i += 256) dst[i] = z[i];
  }
  float m = -INFINITY, ssum = 0.f;
  int arg = 0x7fffffff;
  for (int i0 = lo + tid; i0 < hi; i0 += 16 * 256) {   // one round for the usual 8-way split of 30522
    float v[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int i = i0 + u * 256;
      v[u] = (i < hi) ? __ldcg(z + i) : -INFINITY;
      if (!first && i == static_cast<int>(last)) v[u] = -10000.0f;   // no-repeat (reference :330)
    }
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int i = i0 + u * 256;
      if (v[u] > m) {   // increasing i per thread: keeps the lowest index on exact ties
        ssum = ssum * __expf(m - v[u]) + 1.0f;
        m = v[u];
        arg = i;
      } else if (v[u] != -INFINITY) {
        ssum += __expf(v[u] - m);
      }
    }
  }
  // block reduce of (m, arg, ssum)
  __shared__ float s_m[8], s_s[8];
  __shared__ int s_a[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m_o = __shfl_xor_sync(0xffffffffu, m, o);
    const float s_o = __shfl_xor_sync(0xffffffffu, ssum, o);
    const int a_o = __shfl_xor_sync(0xffffffffu, arg, o);
    const float mn = fmaxf(m, m_o);
    const float sa = (m == -INFINITY) ? 0.f : __expf(m - mn);
    const float sb = (m_o == -INFINITY) ? 0.f : __expf(m_o - mn);
    ssum = ssum * sa + s_o * sb;
    if (m_o > m || (m_o == m && a_o < arg)) arg = a_o;
    m = mn;
  }
  if ((tid & 31) == 0) { s_m[tid >> 5] = m; s_s[tid >> 5] = ssum; s_a[tid >> 5] = arg; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < 8; ++w) {
      const float mn = fmaxf(m, s_m[w]);
      const float sa = (m == -INFINITY) ? 0.f : __expf(m - mn);
      const float sb = (s_m[w] == -INFINITY) ? 0.f : __expf(s_m[w] - mn);
      ssum = ssum * sa + s_s[w] * sb;
      if (s_m[w] > m || (s_m[w] == m && s_a[w] < arg)) arg = s_a[w];
      m = mn;
    }
    p.part_max[row * p.n_split + split] = m;
    p.part_sum[row * p.n_split + split] = ssum;
    p.part_arg[row * p.n_split + split] = arg;
    __threadfence();
    const unsigned int t = atomicAdd(&p.row_ticket[row], 1u);
    if (t == static_cast<unsigned int>(p.n_split) - 1) {
      __threadfence();
      p.row_ticket[row] = 0;
      // combine the slices (slice order = index order, so ">" keeps the lowest index on ties)
      float gm = -INFINITY, gs = 0.f;
      int ga = 0;
      for (int k = 0; k < p.n_split; ++k) {
        const float pm = __ldcg(&p.part_max[row * p.n_split + k]);
        const float ps = __ldcg(&p.part_sum[row * p.n_split + k]);
        const int pa = __ldcg(&p.part_arg[row * p.n_split + k]);
        const float mn = fmaxf(gm, pm);
        const float sa = (gm == -INFINITY) ? 0.f : __expf(gm - mn);
        const float sb = (pm == -INFINITY) ? 0.f : __expf(pm - mn);
        gs = gs * sa + ps * sb;
        if (pm > gm) ga = pa;
        gm = mn;
      }
      const bool row_done = (!first) && (!in_prefix) && (last == p.eos);
      long long tok;
      float lp;
      if (in_prefix) {   // still feeding this row's prefix: the next prefix token, nothing to score
        tok = p.row_prefix[static_cast<long long>(p.row0 + row) * p.row_prefix_stride + cur_len];
        lp = 0.f;
      } else if (row_done) {  // one-hot EOS distribution (reference :347-351): log_softmax gives exactly 0 at EOS
        tok = p.eos;
        lp = 0.f;
      } else {
        tok = ga;
        lp = -logf(gs);  // z[arg] - max - log(sum exp(z - max)) with z[arg] == max
      }
      p.tokens_out[static_cast<long long>(row) * p.max_steps + cur_len] = tok;
      p.logprob_sum[row] += lp;
      long long nxt = tok;
      if (p.forced != nullptr) nxt = p.forced[static_cast<long long>(row) * p.max_steps + cur_len];
      p.next_token[row] = nxt;
      // reference checks `(last_predictions == eos).all()` on the sequence it feeds next
      if (nxt != p.eos) atomicAdd(&st->not_eos, 1);
      __threadfence();
      const unsigned int tr = atomicAdd(&st->ticket, 1u);
      if (tr == static_cast<unsigned int>(p.rows) - 1) {  // last row of this step: advance the loop state
        __threadfence();
        const int not_eos = atomicAdd(&st->not_eos, 0);
        st->ticket = 0;
        st->not_eos = 0;
        st->cur_len = cur_len + 1;
        st->final_len = cur_len + 1;
        st->pos = st->pos + 1;
        st->step = step + 1;
        if (not_eos == 0) {
          st->finished = 1;
          if (step == 0 && p.row_prefix == nullptr) st->empty_caption = 1;
        }
        if (cur_len + 1 >= p.max_steps) st->finished = 1;
        // every CTA of every kernel of this step has passed its wait: recycle the chain counters
        if (p.chain.counters != nullptr)
          for (int k = 0; k < 64; ++k) p.chain.counters[k] = 0;
        __threadfence();
      }
    }
  }
}

// logprobs / num_valid (reference layers/decoder.py:433-438) and EOS padding of the unused tail.
__global__ void greedy_finalize_kernel(long long* tokens_out, const float* logprob_sum, float* logprobs_out, int rows,
                                       int max_steps, int prefix_len, int eos, const StepState* state, const int* row_prefix_lens) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= rows) return;
  if (row_prefix_lens != nullptr) prefix_len = row_prefix_lens[row];
  const int n = state->final_len;
  for (int i = n; i < max_steps; ++i) tokens_out[static_cast<long long>(row) * max_steps + i] = eos;
  int not_eos = 0, has_eos = 0;
  for (int i = 0; i < n; ++i) {
    const long long t = tokens_out[static_cast<long long>(row) * max_steps + i];
    if (t == eos) has_eos = 1; else ++not_eos;
  }
  int num_valid = not_eos + has_eos - prefix_len;
  if (num_valid < 1) num_valid = 1;
  logprobs_out[row] = state->empty_caption ? logprob_sum[row] : logprob_sum[row] / static_cast<float>(num_valid);
}

}  // namespace gitb200
