// Internal C++ declarations shared by the kernels and the C-ABI layer (api.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/cgpt.h"

namespace cgpt {
void count_launch(int n = 1);

int gemm_bf16(const void* A, long long lda, const void* W, long long ldw, int M, int N, int K,
              const cgpt_gemm_epilogue* e, int force_bn, cudaStream_t stream);
int gemm_launch_count();
int gemm_profile_begin();
int gemm_profile_end(float* ms_out, int* mnk_out, int capacity, int* count);
long long total_launch_count();

int noise_patchify(const float* x, const float* eps, uint64_t seed, uint32_t stream_id,
                   uint64_t first_sample, int B, float sigma, const float* mean3,
                   const float* std3, int noise_space, int noise_kind, int img_size, void* out,
                   long long ld_out, const void* dyn, cudaStream_t stream);
int noise_image(const float* x, const float* eps, uint64_t seed, uint32_t stream_id,
                uint64_t first_sample, int B, float sigma, const float* mean3, const float* std3,
                int noise_space, int noise_kind, int channels, int height, int width, float* out,
                cudaStream_t stream);

int answer_labels(const int* ids, int B, int max_new, int ld_ids, int eos_id, const uint64_t* keys,
                  const int* vals, int capacity, int other_label, int* labels, cudaStream_t stream);
uint64_t answer_hash_host(const int* ids, int n, int eos_id);
int argmax_rows(const float* logits, int rows, int cols, long long ld, int suppress_col, int* out_idx,
                float* out_margin, cudaStream_t stream);
int greedy_step(const int* next_idx, int B, int* finished, int* ids_out, int ld, int t, int eos_id,
                int pad_id, int* unfinished_count, cudaStream_t stream);
int label_hist(const int* labels, int B, int num_classes, long long* counts, int* invalid,
               cudaStream_t stream);
int certify_tail(const long long* counts_sel, const long long* counts_est, int num_classes, long long n,
                 double alpha, double sigma, const double* lut, int* out_label, double* out_stats,
                 cudaStream_t stream);
int certify_tail_batch(const long long* counts, int images, int num_classes, long long n, double alpha, double sigma,
                       const double* lut, int* out_label, double* out_stats, cudaStream_t stream);
int label_hist_images(const int* labels, int rows, int per_image, long long first, long long boundary, int num_classes,
                      long long* counts, int* invalid, cudaStream_t stream);
int predict_tail(const long long* counts, int num_classes, double alpha, int* out_label,
                 double* out_stats, cudaStream_t stream);
int ce_loss(const float* logits, long long ld, int rows, int cols, const int* targets, float* token_loss,
            float* mean_count, float label_smoothing, cudaStream_t stream);
int cosine_rows(const float* feats, long long ld, int rows, int D, const float* target, float* scores,
                cudaStream_t stream);
int norm_rows(const void* x, long long ldx, int in_dtype, const float* gamma, const float* beta,
              float eps, int rows, int D, void* out, long long ldo, int out_dtype, int rms,
              int in_row_period, int in_row_stride, int in_row_offset, cudaStream_t stream);
int attention(const cgpt_attn_args* a, cudaStream_t stream);
int attn_umma_supported(const cgpt_attn_args* a);
int attention_umma(const cgpt_attn_args* a, cudaStream_t stream);
int attn_vit_supported(const cgpt_attn_args* a);
int attention_vit(const cgpt_attn_args* a, cudaStream_t stream);
int attn_long_supported(const cgpt_attn_args* a);
int attention_long(const cgpt_attn_args* a, cudaStream_t stream);
int attn_prefill_supported(const cgpt_attn_args* a);
int attention_prefill(const cgpt_attn_args* a, cudaStream_t stream);
int rope_split(void* qkv, long long ld, int rows, int T, int H, int head_dim, int pos0, const float* cos_t,
               const float* sin_t, void* kcache, void* vcache, long long ldc, int cache_rows_per_batch,
               int cache_row0, cudaStream_t stream);
int gather_rows(const void* table, long long ldt, const int* ids, int id_period, int rows, int D, void* out,
                long long ldo, int out_dtype, int remap_period, int remap_stride, int remap_offset, int table_rows,
                cudaStream_t stream);
// fine-tune step (train_ops.cu)
int swiglu_fwd(const void* gu, void* act, long long rows, int inter, cudaStream_t s);
int swiglu_bwd(const void* gu, const void* dact, void* dgu, long long rows, int inter, cudaStream_t s);
int rmsnorm_bwd(const float* x, long long ldx, const float* gamma, const float* dy, long long ldy, float eps, int rows, int D,
                float* dx, long long lddx, int period, int stride, int offset, cudaStream_t s);
int rope_bwd_cast(const float* dqkv, void* out, int rows, int T, int H, int hd, int pos0, const float* cos_t,
                  const float* sin_t, cudaStream_t s);
int attention_bwd(const void* q, long long ldq, const void* kc, const void* vc, long long ldc, int cache_rows, const void* o,
                  long long ldo, const void* dout, long long lddo, float* dqkv, int B, int H, int hd, int Tq, int Tk,
                  float scale, cudaStream_t s);
int ce_grad(const float* logits, long long ld, int rows, int cols, const int* targets, const float* mean_count, void* dlogits,
            long long ldd, float label_smoothing, cudaStream_t s);
int cast_rows_f32_bf16(const float* src, long long lds, void* dst, long long ldd, int rows, int cols, int period, int stride,
                       int offset, cudaStream_t s);
int transpose_bf16(const void* src, long long lds, void* dst, long long ldd, int rows, int cols, cudaStream_t s);
int colsum_bf16(const void* src, long long lds, int rows, int cols, float* out, cudaStream_t s);
int adamw_step(float* p, const float* g, float* m, float* v, void* p_bf16, long long n, float lr, float beta1, float beta2,
               float eps, float wd, int step, float gscale, cudaStream_t s);
}  // namespace cgpt
