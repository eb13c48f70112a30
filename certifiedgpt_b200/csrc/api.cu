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
}  // namespace cgpt

using namespace cgpt;

extern "C" {

const char* cgpt_last_error(void) { return g_err; }
int cgpt_abi_version(void) { return CGPT_ABI_VERSION; }
long long cgpt_launch_count(void) { return g_launches + gemm_launch_count(); }

int cgpt_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, int M, int N, int K,
                   const cgpt_gemm_epilogue* epi, int force_bn, void* stream) {
  return gemm_bf16(A, lda, W, ldw, M, N, K, epi, force_bn, (cudaStream_t)stream);
}

}  // extern "C"
