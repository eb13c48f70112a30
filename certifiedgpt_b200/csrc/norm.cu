// LayerNorm / RMSNorm rows: fp32 statistics, bf16 (or fp32) output feeding the next GEMM.
//   LayerNorm: eva_vit.Block.norm1/norm2 (eps 1e-6, eva_vit.py:162,168,434), ln_vision
//              (fp32 LayerNorm eps 1e-5, base_model.py:281-287), BERT LayerNorm eps 1e-12
//              (Qformer.py:66,104-107,285-289,367-375)
//   RMSNorm  : HF LlamaRMSNorm (x * rsqrt(mean(x^2) + eps) * w), eps from the LLM config
// One warp per row; the row is read with 16-byte loads, statistics are two-pass
// (mean, then centred second moment) like torch's LayerNorm.
#include "common.cuh"
#include "ops.h"

namespace cgpt {

template <bool IN_F32>
__device__ __forceinline__ void load8(const void* base, long long idx, float* v) {
  if (IN_F32) {
    const float4 a = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + idx);
    const float4 b = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + idx + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
    const uint4 a = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(base) + idx);
    v[0] = bf16_lo(a.x); v[1] = bf16_hi(a.x); v[2] = bf16_lo(a.y); v[3] = bf16_hi(a.y);
    v[4] = bf16_lo(a.z); v[5] = bf16_hi(a.z); v[6] = bf16_lo(a.w); v[7] = bf16_hi(a.w);
  }
}

// One warp per row; the row lives in registers between the statistics and the normalisation (a single
// HBM read).  NV = 8-element vectors per lane (D <= 256*NV): 1408 -> 6, 4096 -> 16, 768 -> 3.
template <bool IN_F32, bool OUT_F32, bool RMS, int NV>
__global__ void __launch_bounds__(128) norm_rows_kernel(const void* __restrict__ x, long long ldx,
                                                        const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, float eps,
                                                        int rows, int D, void* __restrict__ out,
                                                        long long ldo, int in_row_period,
                                                        int in_row_stride, int in_row_offset) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  // optional row gather: input row = (row / period) * stride + offset + row % period
  long long in_row = row;
  if (in_row_period > 0)
    in_row = static_cast<long long>(row / in_row_period) * in_row_stride + in_row_offset + (row % in_row_period);
  const long long xoff = in_row * ldx;
  const int nvec = D >> 3;
  float v[NV][8];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int i = lane + 32 * k;
    if (i < nvec) load8<IN_F32>(x, xoff + (i << 3), v[k]);
    else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[k][j] = 0.f;
    }
  }
  float mean = 0.f;
  if (!RMS) {
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k)
#pragma unroll
      for (int j = 0; j < 8; ++j) sum += v[k][j];
    mean = warp_sum(sum) / static_cast<float>(D);
  }
  float sq = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    if (lane + 32 * k < nvec) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float d = v[k][j] - mean; sq += d * d; }
    }
  }
  sq = warp_sum(sq);
  const float rstd = rsqrtf(sq / static_cast<float>(D) + eps);
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int i = lane + 32 * k;
    if (i < nvec) {
      float g[8], b[8], r[8];
      load8<true>(gamma, i << 3, g);
      if (!RMS) load8<true>(beta, i << 3, b);
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = RMS ? (v[k][j] * rstd) * g[j] : (v[k][j] - mean) * rstd * g[j] + b[j];
      if (OUT_F32) {
        float* o = reinterpret_cast<float*>(out) + row * ldo + (i << 3);
        *reinterpret_cast<float4*>(o) = make_float4(r[0], r[1], r[2], r[3]);
        *reinterpret_cast<float4*>(o + 4) = make_float4(r[4], r[5], r[6], r[7]);
      } else {
        __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out) + row * ldo + (i << 3);
        *reinterpret_cast<uint4*>(o) = make_uint4(pack_bf16x2(r[0], r[1]), pack_bf16x2(r[2], r[3]),
                                                  pack_bf16x2(r[4], r[5]), pack_bf16x2(r[6], r[7]));
      }
    }
  }
}

int norm_rows(const void* x, long long ldx, int in_dtype, const float* gamma, const float* beta,
              float eps, int rows, int D, void* out, long long ldo, int out_dtype, int rms,
              int in_row_period, int in_row_stride, int in_row_offset, cudaStream_t stream) {
  CGPT_REQUIRE(rows > 0 && D > 0 && D % 8 == 0, "norm_rows: rows=%d D=%d (D must be a multiple of 8)", rows, D);
  CGPT_REQUIRE(ldx % 8 == 0 && ldo % 8 == 0, "norm_rows: leading dims must be multiples of 8");
  CGPT_REQUIRE(gamma != nullptr && (rms || beta != nullptr), "norm_rows: missing gamma/beta");
  CGPT_REQUIRE(D <= 4096, "norm_rows: D = %d exceeds the register-resident limit (4096)", D);
  const int grid = (rows + 3) / 4;
  const bool inf = in_dtype == CGPT_DT_F32, outf = out_dtype == CGPT_DT_F32;
  const int nv = (D / 8 + 31) / 32;   // vectors per lane
#define LAUNCH4(A, B, C, NVV)                                                                         \
  norm_rows_kernel<A, B, C, NVV><<<grid, 128, 0, stream>>>(x, ldx, gamma, beta, eps, rows, D, out, ldo, \
                                                           in_row_period, in_row_stride, in_row_offset)
#define LAUNCH(A, B, C)                                         \
  do {                                                          \
    if (nv <= 1) LAUNCH4(A, B, C, 1);                           \
    else if (nv <= 3) LAUNCH4(A, B, C, 3);                      \
    else if (nv <= 6) LAUNCH4(A, B, C, 6);                      \
    else LAUNCH4(A, B, C, 16);                                  \
  } while (0)
  if (rms) {
    if (inf) { if (outf) LAUNCH(true, true, true); else LAUNCH(true, false, true); }
    else     { if (outf) LAUNCH(false, true, true); else LAUNCH(false, false, true); }
  } else {
    if (inf) { if (outf) LAUNCH(true, true, false); else LAUNCH(true, false, false); }
    else     { if (outf) LAUNCH(false, true, false); else LAUNCH(false, false, false); }
  }
#undef LAUNCH
#undef LAUNCH4
  CGPT_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace cgpt
