// Internal C++ declarations shared by the kernels and the C-ABI layer (api.cu).
#pragma once
#include <cuda_runtime.h>
#include "../../include/cgpt.h"

namespace cgpt {
void count_launch(int n = 1);

int gemm_bf16(const void* A, long long lda, const void* W, long long ldw, int M, int N, int K,
              const cgpt_gemm_epilogue* e, int force_bn, cudaStream_t stream);
int gemm_launch_count();
}  // namespace cgpt
