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

int noise_patchify(const float* x, const float* eps, uint64_t seed, uint32_t stream_id,
                   uint64_t first_sample, int B, float sigma, const float* mean3,
                   const float* std3, int noise_space, int noise_kind, int img_size, void* out,
                   long long ld_out, cudaStream_t stream);
int noise_image(const float* x, const float* eps, uint64_t seed, uint32_t stream_id,
                uint64_t first_sample, int B, float sigma, const float* mean3, const float* std3,
                int noise_space, int noise_kind, int channels, int height, int width, float* out,
                cudaStream_t stream);

int answer_labels(const int* ids, int B, int max_new, int ld_ids, int eos_id, const uint64_t* keys,
                  const int* vals, int capacity, int other_label, int* labels, cudaStream_t stream);
uint64_t answer_hash_host(const int* ids, int n, int eos_id);
int argmax_rows(const float* logits, int rows, int cols, long long ld, int suppress_col, int* out_idx,
                float* out_margin, cudaStream_t stream);
int label_hist(const int* labels, int B, int num_classes, long long* counts, int* invalid,
               cudaStream_t stream);
int certify_tail(const long long* counts_sel, const long long* counts_est, int num_classes, long long n,
                 double alpha, double sigma, int* out_label, double* out_stats, cudaStream_t stream);
int predict_tail(const long long* counts, int num_classes, double alpha, int* out_label,
                 double* out_stats, cudaStream_t stream);
}  // namespace cgpt
