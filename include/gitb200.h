/* gitb200 -- C ABI of the B200-native GIT captioning engine (libgitb200.so).
 *
 * This is the drop-in boundary for the reference's hot path.  The reference is 100 % Python/PyTorch
 * (there is no FFI of its own), so the "binding a maintainer would add" is a ctypes stub
 * (INTEGRATION.md); each entry point names the reference function it replaces.  Paths are relative to
 * /root/reference/generativeimage2text/.
 *
 * Conventions
 *   - plain C types only; every `dev` pointer is a CUDA device pointer owned by the caller
 *     (e.g. `tensor.data_ptr()`), every `host` pointer is ordinary host memory;
 *   - all work is enqueued on the `stream` argument (a cudaStream_t passed as void*); entry points do
 *     not synchronise unless documented;
 *   - return 0 on success, non-zero on error; `gitb200_last_error` returns a message for the handle
 *     (or for the last failed `gitb200_create` when h == NULL); nothing throws across the ABI;
 *   - one engine per device and per thread of use (the reference model object is not re-entrant
 *     either: layers/decoder.py:991 stores per-call state on the module).
 */
#ifndef GITB200_H_
#define GITB200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gitb200_engine gitb200_engine;

/* Model geometry.  Mirrors what `get_git_model(tokenizer, param)` hard-codes / reads from `param`
 * (model.py:9-61, 63-91): encoder = CLIP ViT-B/16 or ViT-L/14, decoder = 6 x 768 BERT-style layers. */
typedef struct gitb200_config {
  int32_t image_size;      /* param['test_crop_size'] (224)                          model.py:13  */
  int32_t patch;           /* 16 (CLIPViT_B_16) / 14 (CLIPViT_L_14)                   model.py:64-67 */
  int32_t enc_width;       /* 768 / 1024                                                           */
  int32_t enc_layers;      /* 12 / 24                                                              */
  int32_t enc_heads;       /* 12 / 16                                                              */
  int32_t dec_hidden;      /* 768                                                     model.py:17  */
  int32_t dec_layers;      /* 6                                                       model.py:18  */
  int32_t dec_heads;       /* 12                                                      model.py:19  */
  int32_t dec_ffn;         /* 3072                                                    model.py:20  */
  int32_t vocab;           /* 30522                                                   model.py:16  */
  int32_t max_positions;   /* 1024                                                    model.py:21  */
  int32_t num_frames_emb;  /* param['num_image_with_embedding'] or 0                  model.py:59  */
  int32_t sos_id;          /* tokenizer.cls_token_id                                  model.py:54  */
  int32_t eos_id;          /* tokenizer.sep_token_id                                  model.py:35,55 */
} gitb200_config;

/* Search configuration: the two decoders of model.py:27-40. */
enum { GITB200_SEARCH_GREEDY = 0, GITB200_SEARCH_BEAM = 1 };
typedef struct gitb200_search {
  int32_t mode;            /* GREEDY = AutoRegressiveBeamSearch(beam 1, per-node 1)  layers/decoder.py:208-440
                              BEAM   = GeneratorWithBeamSearch                        layers/decoder.py:1056-1341 */
  int32_t max_steps;       /* output length cap incl. start tokens (40 in BASELINE.json; 1024 stock) */
  int32_t beam_size;       /* 4                                                       model.py:38  */
  int32_t per_node_beam;   /* 2                                                       layers/decoder.py:1062 */
  float length_penalty;    /* 0.6                                                     model.py:39  */
} gitb200_search;

enum { GITB200_F32 = 0, GITB200_BF16 = 1, GITB200_I64 = 2 };

/* Replaces: get_git_model (construction)                                            model.py:9-61 */
int gitb200_create(const gitb200_config* cfg, int device, gitb200_engine** out);
void gitb200_destroy(gitb200_engine* h);
const char* gitb200_last_error(const gitb200_engine* h);
/* ABI version of the library (bumped on any signature change). */
int gitb200_abi_version(void);

/* Replaces: torch_common.load_state_dict -> module parameters           torch_common.py:93-145.
 * `ref_key` is the reference state-dict key (e.g. "image_encoder.conv1.weight"); `dev_ptr` an fp32
 * device tensor of `shape[ndim]`.  The engine repacks into its own layouts (bf16 GEMM operands,
 * fused QKV, padded patch kernel); the caller keeps ownership of the source and may free it after
 * gitb200_finalize_weights returns.  Unknown keys return an error; "image_encoder.proj" and
 * "textual.output.weight" (tied) are accepted and ignored. */
int gitb200_set_weight(gitb200_engine* h, const char* ref_key, const void* dev_ptr, const int64_t* shape,
                       int ndim, int dtype, void* stream);
/* Checks that every tensor of the geometry has been provided. Synchronises the stream. */
int gitb200_finalize_weights(gitb200_engine* h, void* stream);
/* Several engines of one geometry on one device (one per batch in flight, see gitb200_generate_async) need only one
 * copy of the parameters: `h` borrows the finalized weight buffers of `src` (which must outlive it) instead of taking
 * its own gitb200_set_weight calls -- the module-level equivalent is calling one nn.Module from several threads.
 * Keeps the decoder weights of all in-flight batches on the same L2 lines. */
int gitb200_share_weights(gitb200_engine* h, gitb200_engine* src);

/* Replaces: CaptioningModel.forward_one image branch = VisualTransformer.forward per frame
 * (+ img_temperal_embedding, token-axis concat)      layers/decoder.py:846-857, layers/CLIP/model.py:240-268.
 * images_dev: fp32 [frames][B,3,H,W] contiguous (frames >= 1; frame f at offset f*B*3*H*W; H = W = image_size unless
 * gitb200_set_input_size says otherwise).
 * feats_out_dev: fp32 [B, frames*L, enc_width] or NULL (kept internally for gitb200_prefill). */
int gitb200_encode(gitb200_engine* h, const float* images_dev, int batch, int frames, float* feats_out_dev,
                   void* stream);

/* Input size of the following gitb200_encode / gitb200_generate* calls when it differs from image_size x image_size
 * (MinMaxResizeForTest inputs, inference.py:29-64): the patch grid becomes (height / patch) x (width / patch) and the
 * positional embedding is re-sampled to it on the device, bicubic, as VisualTransformer.forward does at run time
 *                                                                                  layers/CLIP/model.py:245-251.
 * All images of one call share the size.  Sticky until changed; gitb200_create starts at image_size x image_size. */
int gitb200_set_input_size(gitb200_engine* h, int height, int width);

/* Replaces: visual_projection + the image rows of BertEncoderAsDecoder, computed once (KV cache)
 *                                      layers/decoder.py:535, 92-174; layers/bert/modeling_bert.py:92-334.
 * Uses the features of the last gitb200_encode. vproj_out_dev: fp32 [B, M, 768] or NULL. */
int gitb200_prefill(gitb200_engine* h, int batch, int beam, float* vproj_out_dev, void* stream);

/* Replaces: CaptioningModel.decoding_step (one new token per row)       layers/decoder.py:1013-1054.
 * tokens_dev: int64 [rows] newest token of each row (rows = batch*beam), appended at text position
 * `pos`; beam_idx_dev (int32 [rows], may be NULL) re-orders the text KV cache first
 * (layers/decoder.py:1231).  logits_out_dev: fp32 [rows, vocab] raw last-position logits. */
int gitb200_decode_step(gitb200_engine* h, const int64_t* tokens_dev, const int32_t* beam_idx_dev, int rows,
                        int pos, float* logits_out_dev, void* stream);

/* Replaces: CaptioningModel.forward / infer + decoder.search            layers/decoder.py:838-1011,
 * AutoRegressiveBeamSearch.search :224-440, GeneratorWithBeamSearch.search :1083-1290.
 * images_dev as gitb200_encode.  prefix_dev: int64 [P] start tokens (NULL/0 -> [sos]); reference asserts
 * batch == 1 with a prefix (layers/decoder.py:988).
 * forced_dev (int64 [B, max_steps] or NULL): teacher forcing for parity tests -- at every step the
 * engine records its own choice but feeds forced[:, t] as the next input (greedy only).
 * tokens_out_dev: int64 [B, max_steps] (incl. start tokens; EOS padded);  logprobs_out_dev: fp32 [B];
 * out_len_host: number of valid columns (greedy may stop early, layers/decoder.py:319), written after
 * an internal stream synchronise (the only sync of the call).
 * step_logits_dev: optional fp32 [steps, rows, vocab] dump of raw step logits (parity hook). */
int gitb200_generate(gitb200_engine* h, const float* images_dev, int batch, int frames, const int64_t* prefix_dev,
                     int prefix_len, const gitb200_search* search, const int64_t* forced_dev,
                     int64_t* tokens_out_dev, float* logprobs_out_dev, int32_t* out_len_host,
                     float* step_logits_dev, void* stream);

/* Same as gitb200_generate but with HOST buffers (pinned or pageable): copies the pixels host->device,
 * runs, copies tokens / logprobs back and synchronises.  This is the call a C host makes. */
int gitb200_generate_host(gitb200_engine* h, const float* images_host, int batch, int frames,
                          const int64_t* prefix_host, int prefix_len, const gitb200_search* search,
                          int64_t* tokens_out_host, float* logprobs_out_host, int32_t* out_len_host, void* stream);

/* Asynchronous form of the two calls above (same arguments minus out_len_host): everything is enqueued on `stream`
 * and the call returns; gitb200_generate_finish synchronises that stream and reports the loop length.  One call may
 * be in flight per engine; a host that keeps two engines (two streams) busy overlaps the encoder of batch i+1 with
 * the latency-bound decode loop of batch i (bench.py "pipeline": 2). Output buffers must stay valid until finish. */
int gitb200_generate_async(gitb200_engine* h, const float* images_dev, int batch, int frames, const int64_t* prefix_dev,
                           int prefix_len, const gitb200_search* search, const int64_t* forced_dev,
                           int64_t* tokens_out_dev, float* logprobs_out_dev, float* step_logits_dev, void* stream);
int gitb200_generate_host_async(gitb200_engine* h, const float* images_host, int batch, int frames,
                                const int64_t* prefix_host, int prefix_len, const gitb200_search* search,
                                int64_t* tokens_out_host, float* logprobs_out_host, void* stream);
int gitb200_generate_finish(gitb200_engine* h, int32_t* out_len_host);

/* Measurement hook (bench.py's roofline): device time, by CUDA events on the engine's stream, of the decode loop of the
 * last generate on this engine -- first step launch to last -- with the number of step launches in it and whether each
 * was the single decode_mega_kernel launch.  Waits for that loop to finish.  No reference counterpart. */
int gitb200_last_decode_ms(gitb200_engine* h, float* ms_out, int32_t* steps_out, int32_t* one_kernel_out);

/* Per-row prefixes for the NEXT generate call (question batches; the reference allows one prefix and batch 1 only,
 * layers/decoder.py:985-1006): prefix_dev int64 [rows, stride], lens_dev int32 [rows] (1 <= len <= stride, len < max_steps;
 * tokens past a row's length are ignored).  rows must equal that call's batch; pass prefix_len = 0 to it.  Every row is
 * generated exactly as a batch-1 call with its own prefix would be; tokens_out rows hold prefix + generated tokens.  The
 * buffers must stay valid until the call has finished. */
int gitb200_set_row_prefixes(gitb200_engine* h, const int64_t* prefix_dev, int rows, int stride, const int32_t* lens_dev);

/* Replaces: TrieAutoRegressiveBeamSearch.search + TokenTrie                            trie_decoder.py:27-258
 * (the vocabulary-constrained greedy decoder model.py:42-48 keeps commented out next to the default one).  The trie is
 * passed in CSR form from HOST memory: node n's outgoing edges are [child_begin[n], child_begin[n + 1]), edge e accepts
 * token child_token[e] and leads to node child_node[e]; node 0 is the root.  While a trie is set, GREEDY generate calls
 * raise the log-probs of the allowed next tokens by (max logit - min logit + 1) before the top-1 (:61-62, :141-142),
 * move the cursor (:70, :153) and accumulate the raised value (:163).  Every row of the batch owns a cursor (the
 * reference has one and constrains row 0 only: it is a batch-1 decoder).  Sticky; n_nodes = 0 removes the trie. */
int gitb200_set_trie(gitb200_engine* h, const int32_t* child_begin_host, const int32_t* child_token_host,
                     const int32_t* child_node_host, int n_nodes, int n_edges);

/* Replaces: the do_sample branches of AutoRegressiveBeamSearch.search            layers/decoder.py:260-272, 364-375
 * for the NEXT greedy generate call: row r draws its token at caption length t from softmax(logits / temperature) by an
 * inverse-CDF lookup in index order with uniforms_dev[t * rows + r] (fp32 [steps >= max_steps, rows == batch], device;
 * torch.multinomial's random stream cannot be reproduced, the distribution is the same).  Log-probs as the reference
 * computes them: tempered log-softmax at a row's first decision, un-tempered afterwards.  The buffer must stay valid
 * until the call has finished. */
int gitb200_set_sampling(gitb200_engine* h, const float* uniforms_dev, int steps, int rows, float temperature);

/* Number of kernels the engine launched since creation (bench.py's gpu_launches). */
int64_t gitb200_launch_count(const gitb200_engine* h);
/* Engine switches (defaults in parentheses): use_graph (1) CUDA-graph replay of the decode step, use_pdl (1) programmatic
 * dependent launch inside the step, use_chain (1) flag-ordered decode chain, use_2cta (1) cta_group::2 encoder GEMMs,
 * use_mega (1) greedy decode steps of <= 64 sequences as one persistent kernel (mega_coop (1): launched cooperatively),
 * parity (0) fp32-grade verification mode: every GEMM operand is a (hi, lo) bf16 pair and each product is computed as
 * a_hi w_hi + a_lo w_hi + a_hi w_lo by the same tcgen05 kernel (three K-segments side by side), attention / K/V caches /
 * q, k, v in fp32 -- set it BEFORE gitb200_set_weight (switching it forgets the uploaded weights). */
int gitb200_set_option(gitb200_engine* h, const char* name, int64_t value);

/* ---- single-kernel entry points (unit tests, micro-benchmarks, ncu) -------------------------------- */
/* C = A[M,K] * W[N,K]^T (+bias) (+act: 0 none, 1 QuickGELU, 2 erf-GELU) (+resid fp32 [M,N]).
 * a_dev, w_dev bf16; out fp32 or bf16.  transposed != 0 runs the swap-AB skinny path used by the decode
 * step (a_dev = activations [M<=256,K], w_dev = weight [N,K], out[M,N]); k_splits > 1 (transposed only, raw
 * fp32 out, no bias / act): every split stores its own partial sums, which are added in split order -- bit-reproducible. bn: tile width override (0 = heuristic). */
int gitb200_op_gemm(const void* a_dev, const void* w_dev, const float* bias_dev, const float* resid_dev,
                    void* out_dev, int M, int N, int K, int act, int out_bf16, int transposed, int k_splits,
                    int bn, void* stream);
/* y = LayerNorm(x (+bias) (+resid)) * gamma + beta over the last dim D (768 or 1024), fp32 in,
 * fp32 and/or bf16 out (either may be NULL). */
int gitb200_op_layernorm(const float* x_dev, const float* bias_dev, const float* resid_dev, const float* gamma_dev,
                         const float* beta_dev, float eps, float* out_f32_dev, void* out_bf16_dev, int rows, int D,
                         void* stream);
/* Non-causal multi-head attention over packed bf16 rows: q/k/v [B, S, H*64] with the given row strides
 * (elements) and batch strides; out bf16 [B, S, H*64]. softmax(q k^T / 8) v. */
int gitb200_op_attention(const void* q_dev, const void* k_dev, const void* v_dev, void* out_dev, int B, int S, int H,
                         long long q_row_stride, long long kv_row_stride, long long q_batch_stride,
                         long long kv_batch_stride, long long out_row_stride, long long out_batch_stride,
                         void* stream);

/* ---- test-time image transform on the GPU ----------------------------------------------------------------
 * Replaces: get_image_transform(param)(pil_image)                                       inference.py:111-132
 *   = torchvision Resize(crop, BICUBIC) -> CenterCrop(crop) -> ToTensor -> Normalize(CLIP mean/std), or, with
 *   param['test_respect_ratio_max'], MinMaxResizeForTest                                inference.py:29-64
 * applied to DECODED images (uint8 RGB, HWC -- what PIL's Image.convert('RGB') holds).  Results are bit-identical to
 * the PIL / torchvision pipeline (Pillow's two-pass fixed-point bicubic with antialiasing, libImaging/Resample.c).
 * The caller computes the size rule (it is host arithmetic on two integers, see generativeimage2text_b200/inference.py)
 * and passes one descriptor per image; images of one call may all differ in size. */
typedef struct gitb200_preproc gitb200_preproc;
typedef struct gitb200_image_desc {
  int64_t src_offset;      /* bytes: first pixel of this image inside the packed source buffer                    */
  int32_t src_h, src_w;    /* decoded size                                                                        */
  int32_t resize_h, resize_w; /* size after the bicubic resize (== src: that axis is not resampled)               */
  int32_t crop_top, crop_left; /* window of the resized image that is produced (CenterCrop; 0,0 + full size = none) */
  int32_t out_h, out_w;
  int64_t dst_offset;      /* fp32 elements: the image lands as [3, out_h, out_w] at out_dev + dst_offset          */
} gitb200_image_desc;

int gitb200_preproc_create(int device, gitb200_preproc** out);
void gitb200_preproc_destroy(gitb200_preproc* p);
const char* gitb200_preproc_last_error(const gitb200_preproc* p);
int64_t gitb200_preproc_launch_count(const gitb200_preproc* p);
/* Reserved for future switches (none at present: every call returns non-zero). */
int gitb200_preproc_set_option(gitb200_preproc* p, const char* name, int64_t value);
/* src: packed uint8 RGB images, on the device (src_on_host == 0) or in host memory (pinned or pageable; copied to the
 * device on `stream` first).  mean3 / std3: host float[3].  Work is enqueued on `stream`; a second call on the same
 * handle first waits (on the host) for the previous call's kernels, so use one handle per stream to overlap calls. */
int gitb200_preproc_run(gitb200_preproc* p, const uint8_t* src, int64_t src_bytes, int src_on_host,
                        const gitb200_image_desc* descs_host, int n, const float* mean3, const float* std3,
                        float* out_dev, int64_t out_elems, void* stream);
/* Host-only helper (no GPU needed): the fixed-point weight table of one axis, i.e. Resample.c precompute_coeffs +
 * normalize_coeffs_8bpc for BICUBIC: bounds_out int32 [out_size][2] (first tap, taps), kk_out int32 [out_size][kk_cap]. */
int gitb200_preproc_coeffs(int in_size, int out_size, int32_t* ksize_out, int32_t* bounds_out, int32_t* kk_out, int kk_cap);

/* Debug: copies one decode-step work buffer of the engine ("x", "y" fp32 [rows,768]; "hb", "ctx", "qb" bf16 [rows,768];
 * "ub" bf16 [rows,3072]) to host memory after a device synchronise.  Returns bytes copied or -1. */
long long gitb200_debug_read(gitb200_engine* h, const char* name, void* out_host, long long max_bytes);
/* Debug aid: in-situ timeline of the decode-step kernels. enable != 0 arms it; enable == 0 copies up to
 * max_entries (globaltimer ns, kernel id) pairs to out_host, disarms, and returns the number of entries. */
int gitb200_debug_timeline(int enable, unsigned long long* out_host, int max_entries);

#ifdef __cplusplus
}
#endif
#endif /* GITB200_H_ */
