// extern "C" surface of libcgpt.so (declared in include/cgpt.h).
#include <stdarg.h>
#include <stdio.h>
#include "common.cuh"
#include "ops.h"

namespace cgpt {
static thread_local char g_err[1024] = "";
static long long g_launches = 0;
void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches += n; }
long long total_launch_count() { return g_launches + gemm_launch_count(); }
}  // namespace cgpt

using namespace cgpt;

extern "C" {

const char* cgpt_last_error(void) { return g_err; }
int cgpt_abi_version(void) { return CGPT_ABI_VERSION; }
long long cgpt_launch_count(void) { return g_launches + gemm_launch_count(); }

int cgpt_gemm_profile_begin(void) { return gemm_profile_begin(); }
int cgpt_gemm_profile_end(float* ms_out, int32_t* mnk_out, int capacity, int* count) {
  return gemm_profile_end(ms_out, mnk_out, capacity, count);
}

int cgpt_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, int M, int N, int K,
                   const cgpt_gemm_epilogue* epi, int force_bn, void* stream) {
  return gemm_bf16(A, lda, W, ldw, M, N, K, epi, force_bn, (cudaStream_t)stream);
}

int cgpt_noise_patchify(const float* x, const float* eps, uint64_t seed, uint32_t stream_id,
                        uint64_t first_sample, int B, float sigma, const float* mean3,
                        const float* std3, int noise_space, int noise_kind, int img_size,
                        void* out_patches, int64_t ld_out, void* stream) {
  return noise_patchify(x, eps, seed, stream_id, first_sample, B, sigma, mean3, std3, noise_space,
                        noise_kind, img_size, out_patches, ld_out, nullptr, (cudaStream_t)stream);
}
int cgpt_noise_patchify_dyn(const float* x, const void* dyn_params, int B, const float* mean3,
                            const float* std3, int noise_space, int noise_kind, int img_size,
                            void* out_patches, int64_t ld_out, void* stream) {
  CGPT_REQUIRE(dyn_params != nullptr, "noise_patchify_dyn: dyn_params is null");
  return noise_patchify(x, nullptr, 0, 0, 0, B, 0.f, mean3, std3, noise_space, noise_kind, img_size,
                        out_patches, ld_out, dyn_params, (cudaStream_t)stream);
}
int cgpt_noise_image(const float* x, const float* eps, uint64_t seed, uint32_t stream_id,
                     uint64_t first_sample, int B, float sigma, const float* mean3,
                     const float* std3, int noise_space, int noise_kind, int channels, int height,
                     int width, float* out, void* stream) {
  return noise_image(x, eps, seed, stream_id, first_sample, B, sigma, mean3, std3, noise_space,
                     noise_kind, channels, height, width, out, (cudaStream_t)stream);
}
int cgpt_answer_labels(const int32_t* ids, int B, int max_new, int ld_ids, int eos_id,
                       const uint64_t* table_keys, const int32_t* table_vals, int capacity,
                       int other_label, int32_t* labels, void* stream) {
  return answer_labels(ids, B, max_new, ld_ids, eos_id, table_keys, table_vals, capacity,
                       other_label, labels, (cudaStream_t)stream);
}
uint64_t cgpt_answer_hash(const int32_t* ids, int n, int eos_id) {
  return answer_hash_host(ids, n, eos_id);
}
int cgpt_argmax_rows(const float* logits, int rows, int cols, int64_t ld, int suppress_col,
                     int32_t* out_idx, float* out_margin, void* stream) {
  return argmax_rows(logits, rows, cols, ld, suppress_col, out_idx, out_margin, (cudaStream_t)stream);
}
int cgpt_greedy_step(const int32_t* next_idx, int B, int32_t* finished, int32_t* ids_out, int ld, int t,
                     int eos_id, int pad_id, int32_t* unfinished_count, void* stream) {
  return greedy_step(next_idx, B, finished, ids_out, ld, t, eos_id, pad_id, unfinished_count,
                     (cudaStream_t)stream);
}
int cgpt_label_hist(const int32_t* labels, int B, int num_classes, int64_t* counts,
                    int32_t* invalid, void* stream) {
  return label_hist(labels, B, num_classes, (long long*)counts, invalid, (cudaStream_t)stream);
}
int cgpt_certify_tail(const int64_t* counts_sel, const int64_t* counts_est, int num_classes,
                      int64_t n, double alpha, double sigma, int32_t* out_label,
                      double* out_stats, void* stream) {
  return certify_tail((const long long*)counts_sel, (const long long*)counts_est, num_classes, n,
                      alpha, sigma, nullptr, out_label, out_stats, (cudaStream_t)stream);
}
int cgpt_certify_tail_lut(const int64_t* counts_sel, const int64_t* counts_est, int num_classes,
                          int64_t n, double alpha, double sigma, const double* lut, int32_t* out_label,
                          double* out_stats, void* stream) {
  CGPT_REQUIRE(lut != nullptr, "cgpt_certify_tail_lut: null table");
  return certify_tail((const long long*)counts_sel, (const long long*)counts_est, num_classes, n,
                      alpha, sigma, lut, out_label, out_stats, (cudaStream_t)stream);
}
int cgpt_predict_tail(const int64_t* counts, int num_classes, double alpha, int32_t* out_label,
                      double* out_stats, void* stream) {
  return predict_tail((const long long*)counts, num_classes, alpha, out_label, out_stats,
                      (cudaStream_t)stream);
}

int cgpt_ce_loss(const float* logits, int64_t ld, int rows, int cols, const int32_t* targets, float* token_loss,
                 float* mean_count, void* stream) {
  return ce_loss(logits, ld, rows, cols, targets, token_loss, mean_count, 0.f, (cudaStream_t)stream);
}
int cgpt_ce_loss_smooth(const float* logits, int64_t ld, int rows, int cols, const int32_t* targets, float* token_loss,
                        float* mean_count, float label_smoothing, void* stream) {
  return ce_loss(logits, ld, rows, cols, targets, token_loss, mean_count, label_smoothing, (cudaStream_t)stream);
}

int cgpt_cosine_rows(const float* feats, int64_t ld, int rows, int D, const float* target, float* scores,
                     void* stream) {
  return cosine_rows(feats, ld, rows, D, target, scores, (cudaStream_t)stream);
}

int cgpt_norm_rows(const void* x, int64_t ldx, int in_dtype, const float* gamma, const float* beta,
                   float eps, int rows, int D, void* out, int64_t ldo, int out_dtype, int rms,
                   int in_row_period, int in_row_stride, int in_row_offset, void* stream) {
  return norm_rows(x, ldx, in_dtype, gamma, beta, eps, rows, D, out, ldo, out_dtype, rms, in_row_period,
                   in_row_stride, in_row_offset, (cudaStream_t)stream);
}
int cgpt_attention(const cgpt_attn_args* args, void* stream) { return attention(args, (cudaStream_t)stream); }
int cgpt_rope_split(void* qkv, int64_t ld, int rows, int T, int H, int head_dim, int pos0,
                    const float* cos_table, const float* sin_table, void* kcache, void* vcache,
                    int64_t ld_cache, int cache_rows_per_batch, int cache_row0, void* stream) {
  return rope_split(qkv, ld, rows, T, H, head_dim, pos0, cos_table, sin_table, kcache, vcache, ld_cache,
                    cache_rows_per_batch, cache_row0, (cudaStream_t)stream);
}
int cgpt_gather_rows(const void* table, int64_t ldt, const int32_t* ids, int id_period, int rows, int D,
                     void* out, int64_t ldo, int out_dtype, int remap_period, int remap_stride,
                     int remap_offset, int table_rows, void* stream) {
  return gather_rows(table, ldt, ids, id_period, rows, D, out, ldo, out_dtype, remap_period, remap_stride,
                     remap_offset, table_rows, (cudaStream_t)stream);
}

// ---------------------------------------------------------------- fine-tune step kernels
int cgpt_swiglu_fwd(const void* gu, void* act, int64_t rows, int inter, void* stream) {
  return swiglu_fwd(gu, act, rows, inter, (cudaStream_t)stream);
}
int cgpt_swiglu_bwd(const void* gu, const void* dact, void* dgu, int64_t rows, int inter, void* stream) {
  return swiglu_bwd(gu, dact, dgu, rows, inter, (cudaStream_t)stream);
}
int cgpt_rmsnorm_bwd(const float* x, int64_t ldx, const float* gamma, const float* dy, int64_t ldy, float eps, int rows, int D,
                     float* dx, int64_t lddx, int row_period, int row_stride, int row_offset, void* stream) {
  return rmsnorm_bwd(x, ldx, gamma, dy, ldy, eps, rows, D, dx, lddx, row_period, row_stride, row_offset, (cudaStream_t)stream);
}
int cgpt_rope_bwd_cast(const float* dqkv, void* out, int rows, int T, int H, int head_dim, int pos0, const float* cos_table,
                       const float* sin_table, void* stream) {
  return rope_bwd_cast(dqkv, out, rows, T, H, head_dim, pos0, cos_table, sin_table, (cudaStream_t)stream);
}
int cgpt_attention_bwd(const void* q, int64_t ldq, const void* kcache, const void* vcache, int64_t ld_cache,
                       int cache_rows_per_batch, const void* o, int64_t ldo, const void* dout, int64_t lddo, float* dqkv,
                       int B, int H, int head_dim, int Tq, int Tk, float scale, void* stream) {
  return attention_bwd(q, ldq, kcache, vcache, ld_cache, cache_rows_per_batch, o, ldo, dout, lddo, dqkv, B, H, head_dim, Tq,
                       Tk, scale, (cudaStream_t)stream);
}
int cgpt_ce_grad(const float* logits, int64_t ld, int rows, int cols, const int32_t* targets, const float* mean_count,
                 void* dlogits, int64_t ldd, void* stream) {
  return ce_grad(logits, ld, rows, cols, targets, mean_count, dlogits, ldd, 0.f, (cudaStream_t)stream);
}
int cgpt_ce_grad_smooth(const float* logits, int64_t ld, int rows, int cols, const int32_t* targets, const float* mean_count,
                        void* dlogits, int64_t ldd, float label_smoothing, void* stream) {
  return ce_grad(logits, ld, rows, cols, targets, mean_count, dlogits, ldd, label_smoothing, (cudaStream_t)stream);
}
int cgpt_cast_rows_f32_bf16(const float* src, int64_t lds, void* dst, int64_t ldd, int rows, int cols, int row_period,
                            int row_stride, int row_offset, void* stream) {
  return cast_rows_f32_bf16(src, lds, dst, ldd, rows, cols, row_period, row_stride, row_offset, (cudaStream_t)stream);
}
int cgpt_transpose_bf16(const void* src, int64_t lds, void* dst, int64_t ldd, int rows, int cols, void* stream) {
  return transpose_bf16(src, lds, dst, ldd, rows, cols, (cudaStream_t)stream);
}
int cgpt_colsum_bf16(const void* src, int64_t lds, int rows, int cols, float* out, void* stream) {
  return colsum_bf16(src, lds, rows, cols, out, (cudaStream_t)stream);
}
int cgpt_adamw_step(float* p, const float* g, float* m, float* v, void* p_bf16, int64_t n, float lr, float beta1, float beta2,
                    float eps, float weight_decay, int step, float grad_scale, void* stream) {
  return adamw_step(p, g, m, v, p_bf16, n, lr, beta1, beta2, eps, weight_decay, step, grad_scale, (cudaStream_t)stream);
}

}  // extern "C"
