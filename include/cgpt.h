/* certifiedgpt_b200 C-ABI  (libcgpt.so, sm_100a only)
 *
 * Drop-in boundary for the Monte-Carlo randomized-smoothing hot path of
 * leodesouza/certifiedGPT.  The reference has no FFI (it is pure Python on torch /
 * torch_xla); these entry points are what a Python binding for that path loads with
 * ctypes.  Each declaration cites the reference code it replaces.
 *
 * Conventions
 *   - every function returns int: 0 = ok, <0 = error; cgpt_last_error() gives the
 *     thread-local message.  No C++ exception crosses this boundary.
 *   - all device buffers are caller-owned (e.g. torch tensors' data_ptr()); the library
 *     allocates nothing on the device.  Scratch is passed in explicitly.
 *   - every launch is asynchronous on the cudaStream_t passed as `void* stream`.
 *   - one host thread per GPU / rank; handles are not shared across threads.
 *   - bf16 = __nv_bfloat16 bit pattern; "f32 vectors" (biases, LayerNorm/RMSNorm
 *     parameters, position embeddings) are float.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef CGPT_H_
#define CGPT_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CGPT_ABI_VERSION 3

enum { CGPT_DT_BF16 = 0, CGPT_DT_F32 = 1 };
enum { CGPT_ACT_NONE = 0, CGPT_ACT_GELU = 1, CGPT_ACT_SWIGLU = 2, CGPT_ACT_QUICKGELU = 3 /* x*sigmoid(1.702x), CLIP */ };
enum { CGPT_NOISE_GAUSSIAN = 0, CGPT_NOISE_UNIFORM = 1 };
/* where the noise is added relative to the BLIP Normalize step
 * (processors/base_processor.py:17-34):
 *   NORMALIZED: x is already normalised, out = x + sigma*eps   (smoothing.py:95-97 as written)
 *   PIXEL     : x is in [0,1] pixel space, out = (x + sigma*eps - mean)/std  (north_star) */
enum { CGPT_SPACE_NORMALIZED = 0, CGPT_SPACE_PIXEL = 1 };

const char* cgpt_last_error(void);
int cgpt_abi_version(void);
/* number of kernels this library has launched in this process (bench.py "gpu_launches") */
long long cgpt_launch_count(void);

/* ---------------------------------------------------------------- dense GEMM core
 * out[M,N] = epilogue(A[M,K] . W[N,K]^T), bf16 inputs, fp32 accumulate (tcgen05 + TMEM + TMA).
 * Replaces every nn.Linear / patch Conv2d on the path: eva_vit.py:202 (PatchEmbed.proj),
 * :126-131 (qkv), :151 (proj), :60-64 (fc1/fc2); Qformer.py:128-135,285-289,358-375;
 * minigpt4.py:76-78,141 (llama_proj); HF LlamaForCausalLM q/k/v/o/gate/up/down/lm_head. */
/* optional fused tail of a fused QKV projection with 128-wide heads (HF LlamaAttention: q/k/v_proj, rotary
 * embedding, KV-cache update): N = 3*heads*128; q (rotated) goes to columns [0, heads*128) of `out`, rotated k
 * and plain v go to the KV cache row b*cache_rows_per_batch + cache_row0 + i of row m = b*T + i (position pos0 + i).
 * Replaces the separate cgpt_rope_split pass. */
typedef struct cgpt_gemm_rope {
  int T, heads, pos0;
  const float* cos_table;   /* fp32 [max_pos, 64] */
  const float* sin_table;
  void* kcache;
  void* vcache;
  int64_t ld_cache;
  int cache_rows_per_batch, cache_row0;
} cgpt_gemm_rope;

typedef struct cgpt_gemm_epilogue {
  void* out;            /* [rows, ldo] bf16 or f32                                          */
  int64_t ldo;          /* elements                                                         */
  int out_dtype;        /* CGPT_DT_*                                                        */
  const float* bias;    /* [N] or NULL                                                      */
  const void* resid;    /* residual added after activation, indexed like `out`, or NULL     */
  int64_t ldr;
  int resid_dtype;
  int act;              /* CGPT_ACT_*; SWIGLU: W rows interleaved (gate_j, up_j), out has N/2 cols */
  const float* row_add; /* e.g. pos_embed: adds row_add[(m % row_period) + row_add_offset, n] */
  int64_t ld_row_add;
  int row_period;
  int row_add_offset;
  int remap_stride;     /* >0: out row = (m / row_period) * remap_stride + remap_offset + m % row_period */
  int remap_offset;
  int max_ctas;         /* 0 = one persistent CTA per SM                                    */
  const cgpt_gemm_rope* rope; /* NULL, or the fused rotary + KV-cache-append tail (see above)     */
  /* hm_T > 0: HEAD-MAJOR scatter of a fused q|k|v projection (N = 3 * hm_heads * hm_hd, bf16 out, M % hm_T == 0):
   * element (m, n) goes to out[which][b][h][t][d] with b = m / hm_T, t = m % hm_T, which = n / (heads * hd),
   * h = (n % (heads * hd)) / hd, d = n % hd; each of the three [M / hm_T][heads][hm_T][hd] blocks is dense (ldo unused).
   * The attention kernel then reads one contiguous block per (sample, head) (eva_vit.py:126-131). */
  int hm_T, hm_heads, hm_hd;
} cgpt_gemm_epilogue;

int cgpt_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, int M, int N, int K,
                   const cgpt_gemm_epilogue* epi, int force_bn, void* stream);

/* ---------------------------------------------------------------- K1 fused noise kernel
 * out_patches[b*G*G + py*G + px, c*196 + ky*14 + kx] (bf16, ld_out >= 592, cols 588..591 zero)
 *   = Normalize?( x[c, py*14+ky, px*14+kx] + sigma * eps_b[...] ),  G = img_size/14.
 * eps == NULL: Philox4x32-10 keyed by (seed; element group, first_sample + b, stream_id), so the
 * draw of a sample does not depend on batch size or world size; eps != NULL: injected standard
 * draws [B,3,S,S] fp32 (parity runs).
 * Replaces smoothing.py:95-97 (x.repeat, randn_like*sigma, add), base_processor.py:17-34
 * (Normalize) and the unfold of eva_vit.py:202-209. */
int cgpt_noise_patchify(const float* x, const float* eps, uint64_t seed, uint32_t stream_id,
                        uint64_t first_sample, int B, float sigma, const float* mean3,
                        const float* std3, int noise_space, int noise_kind, int img_size,
                        void* out_patches, int64_t ld_out, void* stream);
/* CUDA-graph friendly variant: (seed, first_sample, stream_id, sigma) are read from the device struct
 * { uint64 seed; uint64 first_sample; uint32 stream_id; float sigma; } at dyn_params, so one captured
 * kernel node serves every batch (the host rewrites the 24-byte struct before each replay) */
int cgpt_noise_patchify_dyn(const float* x, const void* dyn_params, int B, const float* mean3,
                            const float* std3, int noise_space, int noise_kind, int img_size,
                            void* out_patches, int64_t ld_out, void* stream);
/* same draw, NCHW fp32 output for an arbitrary torch base_classifier (smoothing.py:95-97) */
int cgpt_noise_image(const float* x, const float* eps, uint64_t seed, uint32_t stream_id,
                     uint64_t first_sample, int B, float sigma, const float* mean3,
                     const float* std3, int noise_space, int noise_kind, int channels, int height,
                     int width, float* out, void* stream);

/* ---------------------------------------------------------------- labels, histogram, tails
 * cgpt_answer_labels: generated ids [B, ld_ids] -> class id through an open-addressing table of
 * 64-bit FNV-1a hashes of canonical answer token sequences (stop at first eos_id, drop ids
 * 0,1,2); unknown -> other_label.  Token-level restatement of minigpt_base.py:438-446 +
 * agents/minigpt4_eval_agent.py:102 (the reference has no answer->class map).
 * cgpt_answer_hash is the host-side hash used to build the table. */
int cgpt_answer_labels(const int32_t* ids, int B, int max_new, int ld_ids, int eos_id,
                       const uint64_t* table_keys, const int32_t* table_vals, int capacity,
                       int other_label, int32_t* labels, void* stream);
uint64_t cgpt_answer_hash(const int32_t* ids, int n, int eos_id);
/* logits.argmax(1) with lowest-index tie rule (smoothing.py:97) and optional top-2 margin;
 * suppress_col >= 0 masks one column (HF min_length: EOS suppressed on the first new token) */
int cgpt_argmax_rows(const float* logits, int rows, int cols, int64_t ld, int suppress_col,
                     int32_t* out_idx, float* out_margin, void* stream);
/* one step of HF greedy search bookkeeping (generate call, minigpt_base.py:414-427): finished rows
 * emit pad_id, a row finishes on eos_id; *unfinished_count += rows still running (for early exit) */
int cgpt_greedy_step(const int32_t* next_idx, int B, int32_t* finished, int32_t* ids_out, int ld, int t,
                     int eos_id, int pad_id, int32_t* unfinished_count, void* stream);
/* counts[label] += 1  (smoothing.py:98,101-105), no host sync; labels outside [0,num_classes)
 * are not counted and are tallied in *invalid when non-NULL */
int cgpt_label_hist(const int32_t* labels, int B, int num_classes, int64_t* counts,
                    int32_t* invalid, void* stream);
/* smoothing.py:46-56.  out_label[0] = class or -1 (ABSTAIN), out_label[1] = cAHat;
 * out_stats[0] = radius, [1] = pABar, [2] = nA */
int cgpt_certify_tail(const int64_t* counts_sel, const int64_t* counts_est, int num_classes,
                      int64_t n, double alpha, double sigma, int32_t* out_label,
                      double* out_stats, void* stream);
/* The same tail with pABar and Phi^-1(pABar) read from a caller-built table instead of the device bisection:
 * lut = device double[2 * (n + 1)], lut[nA] = _lower_confidence_bound(nA, n, alpha) and lut[n + 1 + nA] =
 * norm.ppf(lut[nA]) for nA = 0..n, tabulated on the host once per (n, alpha) with the very calls the reference makes
 * per image (smoothing.py:55,117: scipy.stats.beta.ppf / norm.ppf).  (label, pABar, radius) are then BIT-IDENTICAL to
 * the reference: the only device arithmetic left is the IEEE fp64 product sigma * lut[n + 1 + nA]. */
int cgpt_certify_tail_lut(const int64_t* counts_sel, const int64_t* counts_est, int num_classes,
                          int64_t n, double alpha, double sigma, const double* lut, int32_t* out_label,
                          double* out_stats, void* stream);
/* smoothing.py:73-79.  out_label[0] = class or -1, [1],[2] = top-2 classes;
 * out_stats[0] = p-value, [1],[2] = top-2 counts */
int cgpt_predict_tail(const int64_t* counts, int num_classes, double alpha, int32_t* out_label,
                      double* out_stats, void* stream);

/* token_loss[r] = logsumexp(logits[r, :]) - logits[r, targets[r]] in fp32, 0 where targets[r] < 0 (ignore_index
 * -100); mean_count (nullable) = {mean over the counted rows, their number} by a fixed-order reduction.
 * Plain CrossEntropyLoss(reduction='mean') of the training / validation forward (modeling_llama.py:101-123); the reference's
 * subclass adds label_smoothing=0.1 there (:107), which these kernels do not apply yet. */
int cgpt_ce_loss(const float* logits, int64_t ld, int rows, int cols, const int32_t* targets, float* token_loss,
                 float* mean_count, void* stream);
/* the reference's loss: token_loss[r] = logsumexp(z) - (1 - eps) * z[target] - eps * mean(z), eps = label_smoothing
 * (CrossEntropyLoss(label_smoothing=0.1), modeling_llama.py:107) */
int cgpt_ce_loss_smooth(const float* logits, int64_t ld, int rows, int cols, const int32_t* targets, float* token_loss,
                        float* mean_count, float label_smoothing, void* stream);

/* scores[r] = cos(feats[r, :], target) in fp32 (one warp per row): the CLIP feature cosine of the black-box
 * attack loop (BASELINE.json configs[4]; README.md:62-64 - the reference ships no code for it) */
int cgpt_cosine_rows(const float* feats, int64_t ld, int rows, int D, const float* target, float* scores,
                     void* stream);

/* ---------------------------------------------------------------- norms
 * LayerNorm (rms = 0) or RMSNorm (rms = 1) over rows of width D; fp32 statistics.
 * Replaces nn.LayerNorm in eva_vit.Block (eva_vit.py:162,168; eps 1e-6), ln_vision
 * (base_model.py:281-287; eps 1e-5), BERT LayerNorms (Qformer.py:66,285-289,367-375; eps 1e-12)
 * and HF LlamaRMSNorm.  in_row_period > 0 gathers input row
 * (r / period) * in_row_stride + in_row_offset + r % period (e.g. last token of each sample). */
int cgpt_norm_rows(const void* x, int64_t ldx, int in_dtype, const float* gamma, const float* beta,
                   float eps, int rows, int D, void* out, int64_t ldo, int out_dtype, int rms,
                   int in_row_period, int in_row_stride, int in_row_offset, void* stream);

/* ---------------------------------------------------------------- attention
 * softmax(scale * Q K^T [+ causal mask]) V per (batch, head), bf16 in/out, fp32 softmax.
 * Head h occupies columns [h*head_dim, (h+1)*head_dim) of the q/k/v/o rows.  Key rows
 * [0, P) come from the batch-invariant prefix buffers kp/vp (shared prompt-prefix KV cache),
 * rows [P, Tk) from this batch's k/v rows.  causal: query i sees keys <= i + (Tk - Tq).
 * Replaces eva_vit.Attention.forward (eva_vit.py:133-150), BertSelfAttention.forward
 * (Qformer.py:195-275) and HF LlamaAttention inside generate (minigpt_base.py:414-427). */
typedef struct cgpt_attn_args {
  const void* q; int64_t ldq; int q_rows_per_batch;
  const void* k; const void* v; int64_t ldk; int64_t ldv; int kv_rows_per_batch;
  const void* kp; const void* vp; int64_t ldkp; int64_t ldvp; int P;
  void* o; int64_t ldo;
  int B, H, Tq, Tk, head_dim;
  float scale;
  int causal;
  int decode_kernel; /* kernel selector: 0 = auto (tcgen05 one-shot kernel when the shape allows, else the
                        mma.sync flash kernel), 1 = single-token KV-cache kernel (Tq == 1), 2 = force flash */
  int head_major;    /* 1: q, k, v are HEAD-MAJOR [B][H][T][head_dim] blocks (what the fused-QKV GEMM writes with
                        cgpt_gemm_epilogue.hm_T; ldq / ldk / ldv unused, Tq == Tk == rows per batch): served by the
                        pipelined tcgen05 kernels, non-causal, 64 < hd <= 128: csrc/attn_vit.cu up to 256 (+1 cls) keys
                        (224 px ViT), csrc/attn_long.cu beyond (448 px ViT, T = 1025: key tiles, two-pass softmax) */
} cgpt_attn_args;
int cgpt_attention(const cgpt_attn_args* args, void* stream);

/* ---------------------------------------------------------------- language-head helpers
 * HF Llama rotary embedding (rotate_half) on the q,k parts of a fused [rows, 3*H*hd] QKV
 * buffer; q in place, rotated k and v appended to the KV cache at row
 * b*cache_rows_per_batch + cache_row0 + i (row m = b*T + i, position pos0 + i).
 * cos/sin tables: fp32 [max_pos, hd/2]. */
int cgpt_rope_split(void* qkv, int64_t ld, int rows, int T, int H, int head_dim, int pos0,
                    const float* cos_table, const float* sin_table, void* kcache, void* vcache,
                    int64_t ld_cache, int cache_rows_per_batch, int cache_row0, void* stream);
/* out[remap(r)] = table[ids ? ids[r % id_period] : r % id_period]  (embed_tokens gather and
 * row broadcast: minigpt_base.py:75-89,367-372,399-412; query_tokens.expand minigpt4.py:133).
 * table_rows > 0: ids outside [0, table_rows) read row 0 instead of an address outside the table. */
int cgpt_gather_rows(const void* table, int64_t ldt, const int32_t* ids, int id_period, int rows, int D,
                     void* out, int64_t ldo, int out_dtype, int remap_period, int remap_stride,
                     int remap_offset, int table_rows, void* stream);


/* ---------------------------------------------------------------- GEMM timing hook (bench.py roofline)
 * begin: from now on every cgpt_gemm_bf16 / engine GEMM launch (eager launches only, not graph replays) is
 * bracketed by a CUDA-event pair on its stream.  end: synchronises, writes up to `capacity` records
 * (ms_out[i], mnk_out[3*i..3*i+2] = M,N,K) and the total number of records to *count, stops recording. */
int cgpt_gemm_profile_begin(void);
int cgpt_gemm_profile_end(float* ms_out, int32_t* mnk_out, int capacity, int* count);

/* ---------------------------------------------------------------- fine-tune step: backward + optimiser kernels
 * MiniGPT4FineTuneAgent.train (agents/minigpt4_finetune_agent.py:149-195): loss.backward() through the frozen Llama to
 * llama_proj (the only trainable module: base_model.py:162-172,238-240; minigpt4.py:76-78,111-117) and an AdamW step.
 * The data-gradient GEMMs are cgpt_gemm_bf16 on transposed weight copies; these are the non-GEMM pieces (fp32 math,
 * fixed summation order).  Host orchestration: certifiedgpt_b200/train.py. */
/* SwiGLU on a fused gate/up GEMM output gu [rows, 2*inter] bf16, (gate_j, up_j) interleaved: act = silu(gate) * up */
int cgpt_swiglu_fwd(const void* gu, void* act, int64_t rows, int inter, void* stream);
int cgpt_swiglu_bwd(const void* gu, const void* dact, void* dgu, int64_t rows, int inter, void* stream);
/* LlamaRMSNorm backward, accumulated into dx (the residual gradient): dx[map(r)] += d/dx (x r gamma) . dy[r];
 * row_period > 0: map(r) = (r / period) * stride + offset + r % period (rows the final norm was applied to) */
int cgpt_rmsnorm_bwd(const float* x, int64_t ldx, const float* gamma, const float* dy, int64_t ldy, float eps, int rows, int D,
                     float* dx, int64_t lddx, int row_period, int row_stride, int row_offset, void* stream);
/* transpose of HF's rotary rotation on the q, k parts of dqkv f32 [rows, 3*H*hd] (v copied), cast to bf16 */
int cgpt_rope_bwd_cast(const float* dqkv, void* out, int rows, int T, int H, int head_dim, int pos0, const float* cos_table,
                       const float* sin_table, void* stream);
/* causal attention backward for short sequences (Tq, Tk <= ~128): one CTA per (sample, head); q rotated queries,
 * keys / values = rows [0, Tk) of the sample's KV cache (shared-prefix rows first); dqkv f32 [B*Tq, 3*H*hd] receives dq and
 * the dk / dv of the sample's own rows (prefix keys have no gradient consumer) */
int cgpt_attention_bwd(const void* q, int64_t ldq, const void* kcache, const void* vcache, int64_t ld_cache,
                       int cache_rows_per_batch, const void* o, int64_t ldo, const void* dout, int64_t lddo, float* dqkv,
                       int B, int H, int head_dim, int Tq, int Tk, float scale, void* stream);
/* dlogits bf16 = (softmax(logits) - onehot(target)) / count, zero rows where target < 0; mean_count from cgpt_ce_loss */
int cgpt_ce_grad(const float* logits, int64_t ld, int rows, int cols, const int32_t* targets, const float* mean_count,
                 void* dlogits, int64_t ldd, void* stream);
/* dlogits = (softmax - (1 - eps) * onehot - eps / cols) / count: gradient of cgpt_ce_loss_smooth's mean */
int cgpt_ce_grad_smooth(const float* logits, int64_t ld, int rows, int cols, const int32_t* targets, const float* mean_count,
                        void* dlogits, int64_t ldd, float label_smoothing, void* stream);
int cgpt_cast_rows_f32_bf16(const float* src, int64_t lds, void* dst, int64_t ldd, int rows, int cols, int row_period,
                            int row_stride, int row_offset, void* stream);
int cgpt_transpose_bf16(const void* src, int64_t lds, void* dst, int64_t ldd, int rows, int cols, void* stream);
int cgpt_colsum_bf16(const void* src, int64_t lds, int rows, int cols, float* out, void* stream);
/* torch.optim.AdamW step on fp32 master weights (decoupled weight decay), refreshing the bf16 copy the GEMMs read */
int cgpt_adamw_step(float* p, const float* g, float* m, float* v, void* p_bf16, int64_t n, float lr, float beta1, float beta2,
                    float eps, float weight_decay, int step, float grad_scale, void* stream);

/* ================================================================= native engine
 * The whole MiniGPT-4 noisy-sample classifier behind one handle: host-side
 * C++ orchestration of the kernels above, CUDA-graph replay per batch size, no Python in the loop.
 *   cgpt_create -> cgpt_bind_weight (x every packed tensor) -> cgpt_set_prompt -> cgpt_set_answer_table
 *   -> cgpt_workspace_bytes / cgpt_bind_workspace (caller-owned device memory; the library allocates
 *   nothing on the device) -> cgpt_sample_noise / cgpt_certify / cgpt_predict, or the per-subsystem
 *   entry points cgpt_vit_forward / cgpt_qformer_forward / cgpt_llm_prefill_decode.
 * Replaces: Smooth._sample_noise / certify / predict (randomized_smoothing/smoothing.py:29-117),
 * MiniGPT4.encode_img (minigpt4.py:121-149) and MiniGPTBase.generate (minigpt_base.py:374-448). */
typedef struct cgpt_model_config {
  /* EVA ViT (eva_vit.py:425-437) */
  int img_size, vit_dim, vit_depth, vit_heads, vit_mlp;
  float vit_eps, ln_vision_eps;
  /* Q-Former (minigpt4.py:90-119) */
  int qf_hidden, qf_layers, qf_heads, qf_inter, qf_queries, qf_cross_freq;
  float qf_eps;
  /* Llama (HF LlamaConfig) */
  int llm_hidden, llm_layers, llm_heads, llm_inter, llm_vocab;
  float llm_rms_eps;
  int eos_id, pad_id;
  /* generation + adapter (minigpt_base.py:374-448) */
  int n_prefix, n_suffix;          /* prompt token counts before the image / after it        */
  int max_new_tokens, min_length;
  int num_classes;                 /* unknown answers map to num_classes - 1 ("other")       */
  int early_exit;                  /* 1: stop decoding when every row has emitted EOS (HF)   */
  int use_graphs;                  /* 1: CUDA-graph replay of the per-batch kernel sequence  */
} cgpt_model_config;

typedef struct cgpt_engine* cgpt_handle;

/* how the noise of one certify/predict call is drawn (smoothing.py:95-97; see cgpt_noise_patchify) */
typedef struct cgpt_noise_spec {
  uint64_t seed;
  uint32_t stream_id;              /* image id (Philox key)                                   */
  float sigma;
  float mean[3], std[3];           /* BLIP Normalize constants (used when noise_space = PIXEL) */
  int noise_space, noise_kind;
  const float* eps;                /* NULL = Philox; else injected standard draws for global sample 0, 1, ...
                                      ([n_total, 3, S, S] fp32, device) */
} cgpt_noise_spec;

int cgpt_create(const cgpt_model_config* cfg, cgpt_handle* out);
int cgpt_destroy(cgpt_handle h);
/* Binds one packed tensor by name (device pointer, caller keeps it alive).  bf16 matrices are K-major
 * [out, in] (nn.Linear layout); names and packing: certifiedgpt_b200/engine.py::_pack. */
int cgpt_bind_weight(cgpt_handle h, const char* name, const void* ptr, int64_t rows, int64_t cols, int dtype);
/* The question of the NEXT calls (the tokens after the image: "</Img> {question} [/INST]", minigpt_base.py:75-89):
 * n_suffix <= cgpt_model_config.n_suffix, which sizes the workspace and the KV cache (the maximum question length).
 * suffix_ids: device int32 [n_suffix], caller-owned.  Every dataset item has its own question
 * (datasets/datasets/vqav2_dataset.py:19-166): the certify / predict agents call this once per item.  Captured graphs
 * are keyed by (n_suffix, suffix_ids): rewrite one device buffer in place to replay them. */
int cgpt_set_question(cgpt_handle h, const int32_t* suffix_ids, int n_suffix);
/* prompt token ids around the image (minigpt_base.py:75-89): device int32 arrays of n_prefix / n_suffix ids */
int cgpt_set_prompt(cgpt_handle h, const int32_t* prefix_ids, const int32_t* suffix_ids);
int cgpt_set_answer_table(cgpt_handle h, const uint64_t* table_keys, const int32_t* table_vals, int capacity);
/* device bytes needed for batches of up to max_batch samples (encoder_only != 0: no LLM buffers / KV cache) */
int cgpt_workspace_bytes(cgpt_handle h, int max_batch, int encoder_only, int64_t* bytes);
/* carves the workspace, computes the shared prompt-prefix K/V once and replicates it into the KV cache */
int cgpt_bind_workspace(cgpt_handle h, void* workspace, int64_t bytes, int max_batch, int encoder_only,
                        void* stream);

/* (2) EVA ViT-g + ln_vision: patches bf16 [B*G*G, 592] -> image tokens bf16 [B*T, vit_dim] */
int cgpt_vit_forward(cgpt_handle h, const void* patches, int B, void* out_tokens, void* stream);
/* (2) Q-Former (+ optional llama_proj): tokens bf16 [B*T, vit_dim] -> queries bf16 [B*32, qf_hidden];
 * out_llm_embeds (nullable) receives llama_proj(queries) bf16 [B*32, llm_hidden] (encode_img's inputs_llama) */
int cgpt_qformer_forward(cgpt_handle h, const void* tokens, int B, void* out_queries, void* out_llm_embeds,
                         void* stream);
/* (3) llama_proj + prompt assembly + prefill over the shared prefix KV + greedy decode:
 * queries bf16 [B*32, qf_hidden] -> out_ids int32 [B, max_new_tokens], out_top2_margin f32 [B, max_new_tokens]
 * (nullable), *out_steps (host, nullable) = decode steps actually run.  Synchronises the stream between
 * steps only when early_exit is set. */
int cgpt_llm_prefill_decode(cgpt_handle h, const void* queries, int B, int32_t* out_ids, float* out_top2_margin,
                            int* out_steps, void* stream);
/* Teacher-forced language-model loss of the fine-tune / validation forward (MiniGPTBase.forward,
 * minigpt_base.py:323-362; loss modeling_llama.py:101-123): B DIFFERENT images, already noised and patchified by
 * the caller (patches bf16 [B*G*G, 592]; agents/minigpt4_finetune_agent.py:142-148 adds uniform noise), go through
 * ViT -> Q-Former -> llama_proj; the sequence [prefix | image | suffix | answer] runs through the frozen Llama and the
 * answer tokens are scored.  answer_ids: device int32 [B, na], -100 = padding (ignored); na <= max_new_tokens.
 * out_token_loss: device f32 [B*na]; out_mean_count: device f32 [2] = {mean loss over scored tokens, their count}.
 * Forward only: the backward pass to the llama_proj gradient is orchestrated by certifiedgpt_b200/train.py over the
 * backward / optimiser kernels declared above (cgpt_attention_bwd, cgpt_rmsnorm_bwd, cgpt_ce_grad, cgpt_adamw_step). */
int cgpt_lm_loss(cgpt_handle h, const void* patches, int B, const int32_t* answer_ids, int na, float* out_token_loss,
                 float* out_mean_count, void* stream);
/* one batch of the hot loop: labels[b] = class of f(x + sigma * eps_{first_sample + b}), b < B.
 * x: fp32 [3,S,S] device pointer.  labels: device int32 [B]. */
int cgpt_noisy_labels(cgpt_handle h, const float* x, const cgpt_noise_spec* noise, uint64_t first_sample, int B,
                      int32_t* labels, void* stream);
/* Smooth._sample_noise (smoothing.py:81-99) over the global sample range [base, base+num), this rank's
 * contiguous slice of it when world > 1.  counts: device int64 [nvec * num_classes], zeroed here; with
 * split >= 0 samples [base, base+split) are counted into vector 0 and the rest into vector 1 (nvec = 2).
 * x may be a host or a device pointer.  comm (nullable): all-reduces the counts (cgpt_comm_init).
 * No host synchronisation unless early_exit is set. */
int cgpt_sample_noise(cgpt_handle h, const float* x, const cgpt_noise_spec* noise, int64_t base, int64_t num,
                      int batch_size, int64_t split, int rank, int world, void* comm, int64_t* counts,
                      int32_t* invalid, void* stream);
/* Smooth.certify (smoothing.py:29-56): one fused pass over [0, n0+n), device tail, one D2H read.
 * x: host or device fp32 [3,S,S].  out_label (host): class or -1 (ABSTAIN); out_radius (host);
 * out_detail (host, nullable) double[3] = {cAHat, pABar, nA}. */
int cgpt_certify(cgpt_handle h, const float* x, const cgpt_noise_spec* noise, int64_t n0, int64_t n, double alpha,
                 int batch_size, int rank, int world, void* comm, int* out_label, double* out_radius,
                 double* out_detail, void* stream);
/* Smooth.certify of K images in shared passes (BASELINE.json configs[2]: a 64-image subset whose draws are sharded over
 * 8 GPUs): every pass holds this rank's next draws of ALL K images, so the per-rank batch stays large (8 images x 138 draws
 * instead of 138).  xs: host array of K pointers to fp32 [3,S,S] images (host or device memory).  Image k is drawn from
 * Philox stream noise->stream_id + k, draw i = global sample index i: per-image counts, labels and radii equal K separate
 * cgpt_certify calls bit for bit, whatever K, batch_size and world are.  One all-reduce of K * 2 * num_classes int64
 * counts, K tails on the device, one D2H read.  Needs cgpt_set_option(h, "max_images", >= K) before the workspace is bound.
 * out_labels / out_radii: host [K]; out_detail (host, nullable): [K][3] = {cAHat, pABar, nA}. */
int cgpt_certify_batch(cgpt_handle h, const float* const* xs, int K, const cgpt_noise_spec* noise, int64_t n0, int64_t n,
                       double alpha, int batch_size, int rank, int world, void* comm, int* out_labels, double* out_radii,
                       double* out_detail, void* stream);
/* Bind the (n, alpha) table of cgpt_certify_tail_lut to the handle (caller-owned device memory, like the weights;
 * NULL unbinds): cgpt_certify calls with exactly this n and alpha take pABar / radius from it (bit-identical to the
 * reference's SciPy tail), every other call uses the device bisection. */
int cgpt_set_radius_lut(cgpt_handle h, int64_t n, double alpha, const double* lut);
/* Smooth.predict (smoothing.py:58-79); out_detail (host, nullable) double[1] = {p-value} */
int cgpt_predict(cgpt_handle h, const float* x, const cgpt_noise_spec* noise, int64_t n, double alpha,
                 int batch_size, int rank, int world, void* comm, int* out_label, double* out_detail,
                 void* stream);
/* device count vectors of the last cgpt_certify / cgpt_predict (int64 [2 * num_classes]; diagnostics) */
int cgpt_last_counts(cgpt_handle h, const int64_t** counts);
/* decode steps run by the last batch (< max_new_tokens when early_exit stopped the loop), or -1 */
int cgpt_last_decode_steps(cgpt_handle h);
/* run-time switches: "use_graphs" (0 = launch every kernel eagerly, e.g. for per-kernel timing),
 * "early_exit", "label_smoothing_permille" (cgpt_lm_loss: 100 = the reference's label_smoothing=0.1),
 * "max_images" (images per pass of cgpt_certify_batch; changing it unbinds the workspace: size and bind it again) */
int cgpt_set_option(cgpt_handle h, const char* key, int value);

/* ---------------------------------------------------------------- the one collective on the path
 * int64 label-count all-reduce over NCCL (NVLink / NVSwitch).  The communicator is created from a
 * 128-byte ncclUniqueId that rank 0 obtains and the caller broadcasts (e.g. torch.distributed).
 * libnccl.so.2 is resolved at run time (dlopen); without it these calls fail. */
int cgpt_comm_unique_id(void* id128);
int cgpt_comm_init(const void* id128, int rank, int world, void** comm);
int cgpt_comm_destroy(void* comm);
int cgpt_allreduce_counts(int64_t* counts, int n, void* comm, void* stream);
/* fp32 sum over the ranks: the llama_proj gradient of the fine-tune step (xm.reduce_gradients,
 * agents/minigpt4_finetune_agent.py:174), 3.1 M values */
int cgpt_allreduce_f32(float* values, int64_t n, void* comm, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CGPT_H_ */
