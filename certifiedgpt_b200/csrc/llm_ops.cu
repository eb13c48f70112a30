// Small memory-bound helpers around the language head:
//   rope_split  : HF Llama rotary embedding (rotate_half convention) applied to the q and k parts
//                 of a fused QKV GEMM output; rotated k and plain v are appended to the KV cache
//   gather_rows : embedding lookup / row broadcast with optional row remap and fp32 conversion
//                 (embed_tokens + concat of [prefix | image | question] embeddings,
//                 minigpt_base.py:75-89,367-372,399-412; query_tokens.expand, minigpt4.py:133)
#include "common.cuh"
#include "ops.h"

namespace cgpt {

// qkv: [rows, 3*H*HD] bf16, row m = b*T + i; position = pos0 + i
// q is rotated in place; k (rotated) and v go to cache row b*cache_rows_per_batch + cache_row0 + i
template <int HD>
__global__ void __launch_bounds__(256) rope_split_kernel(__nv_bfloat16* __restrict__ qkv, long long ld,
                                                         int T, int H, int pos0,
                                                         const float* __restrict__ cos_t,
                                                         const float* __restrict__ sin_t,
                                                         __nv_bfloat16* __restrict__ kcache,
                                                         __nv_bfloat16* __restrict__ vcache, long long ldc,
                                                         int cache_rows_per_batch, int cache_row0) {
  constexpr int HALF = HD / 2;
  constexpr int CH = HALF / 8;  // 16-byte chunks per half head
  const int m = blockIdx.x;
  const int b = m / T, i = m - b * T;
  const int pos = pos0 + i;
  __nv_bfloat16* row = qkv + static_cast<long long>(m) * ld;
  const long long crow = (static_cast<long long>(b) * cache_rows_per_batch + cache_row0 + i) * ldc;
  const int D = H * HD;
  const float* cs = cos_t + static_cast<long long>(pos) * HALF;
  const float* sn = sin_t + static_cast<long long>(pos) * HALF;
  for (int w = threadIdx.x; w < 2 * H * CH; w += blockDim.x) {
    const int part = w / (H * CH);  // 0 = q, 1 = k
    const int rem = w - part * H * CH;
    const int h = rem / CH, c = rem - h * CH;
    __nv_bfloat16* src = row + part * D + h * HD + c * 8;
    const uint4 lo = *reinterpret_cast<const uint4*>(src);
    const uint4 hi = *reinterpret_cast<const uint4*>(src + HALF);
    float x1[8] = {bf16_lo(lo.x), bf16_hi(lo.x), bf16_lo(lo.y), bf16_hi(lo.y),
                   bf16_lo(lo.z), bf16_hi(lo.z), bf16_lo(lo.w), bf16_hi(lo.w)};
    float x2[8] = {bf16_lo(hi.x), bf16_hi(hi.x), bf16_lo(hi.y), bf16_hi(hi.y),
                   bf16_lo(hi.z), bf16_hi(hi.z), bf16_lo(hi.w), bf16_hi(hi.w)};
    float o1[8], o2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float cc = cs[c * 8 + j], ss = sn[c * 8 + j];
      o1[j] = x1[j] * cc - x2[j] * ss;  // x*cos + rotate_half(x)*sin, first half: -x2
      o2[j] = x2[j] * cc + x1[j] * ss;  // second half: +x1
    }
    const uint4 r1 = make_uint4(pack_bf16x2(o1[0], o1[1]), pack_bf16x2(o1[2], o1[3]),
                                pack_bf16x2(o1[4], o1[5]), pack_bf16x2(o1[6], o1[7]));
    const uint4 r2 = make_uint4(pack_bf16x2(o2[0], o2[1]), pack_bf16x2(o2[2], o2[3]),
                                pack_bf16x2(o2[4], o2[5]), pack_bf16x2(o2[6], o2[7]));
    __nv_bfloat16* dst = part == 0 ? src : kcache + crow + h * HD + c * 8;
    *reinterpret_cast<uint4*>(dst) = r1;
    *reinterpret_cast<uint4*>(dst + HALF) = r2;
  }
  const uint4* vsrc = reinterpret_cast<const uint4*>(row + 2 * D);
  uint4* vdst = reinterpret_cast<uint4*>(vcache + crow);
  for (int w = threadIdx.x; w < D / 8; w += blockDim.x) vdst[w] = vsrc[w];
}

int rope_split(void* qkv, long long ld, int rows, int T, int H, int head_dim, int pos0, const float* cos_t,
               const float* sin_t, void* kcache, void* vcache, long long ldc, int cache_rows_per_batch,
               int cache_row0, cudaStream_t stream) {
  CGPT_REQUIRE(rows > 0 && T > 0 && rows % T == 0, "rope_split: rows=%d must be a multiple of T=%d", rows, T);
  CGPT_REQUIRE(ld % 8 == 0 && ldc % 8 == 0, "rope_split: leading dims must be multiples of 8");
  CGPT_REQUIRE(cache_row0 + T <= cache_rows_per_batch, "rope_split: cache overflow (%d + %d > %d)", cache_row0, T,
               cache_rows_per_batch);
#define RL(HDV)                                                                                                  \
  rope_split_kernel<HDV><<<rows, 256, 0, stream>>>((__nv_bfloat16*)qkv, ld, T, H, pos0, cos_t, sin_t,            \
                                                   (__nv_bfloat16*)kcache, (__nv_bfloat16*)vcache, ldc,           \
                                                   cache_rows_per_batch, cache_row0)
  if (head_dim == 128) RL(128);
  else if (head_dim == 64) RL(64);
  else if (head_dim == 32) RL(32);
  else if (head_dim == 16) RL(16);
  else CGPT_REQUIRE(false, "rope_split: head_dim %d unsupported", head_dim);
#undef RL
  CGPT_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

// out[remap(r), :] = table[ids ? ids[r % id_period] : r % id_period, :]   (bf16 table)
template <bool OUT_F32>
__global__ void __launch_bounds__(128) gather_rows_kernel(const __nv_bfloat16* __restrict__ table, long long ldt,
                                                          const int* __restrict__ ids, int id_period, int D,
                                                          void* __restrict__ out, long long ldo,
                                                          int remap_period, int remap_stride, int remap_offset,
                                                          int table_rows) {
  const int r = blockIdx.x;
  const int sel = r % id_period;
  long long src_row = ids ? ids[sel] : sel;
  // table_rows > 0: ids outside the table read row 0 instead of a wild address (an id can only leave the table
  // through a poisoned upstream value, e.g. an all-NaN logits row; the caller validates user-supplied ids)
  if (table_rows > 0 && (src_row < 0 || src_row >= table_rows)) src_row = 0;
  long long orow = r;
  if (remap_period > 0)
    orow = static_cast<long long>(r / remap_period) * remap_stride + remap_offset + (r % remap_period);
  const uint4* src = reinterpret_cast<const uint4*>(table + src_row * ldt);
  for (int c = threadIdx.x; c < D / 8; c += blockDim.x) {
    const uint4 v = src[c];
    if (OUT_F32) {
      float* o = reinterpret_cast<float*>(out) + orow * ldo + c * 8;
      *reinterpret_cast<float4*>(o) = make_float4(bf16_lo(v.x), bf16_hi(v.x), bf16_lo(v.y), bf16_hi(v.y));
      *reinterpret_cast<float4*>(o + 4) = make_float4(bf16_lo(v.z), bf16_hi(v.z), bf16_lo(v.w), bf16_hi(v.w));
    } else {
      *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(out) + orow * ldo + c * 8) = v;
    }
  }
}

int gather_rows(const void* table, long long ldt, const int* ids, int id_period, int rows, int D, void* out,
                long long ldo, int out_dtype, int remap_period, int remap_stride, int remap_offset, int table_rows,
                cudaStream_t stream) {
  CGPT_REQUIRE(rows > 0 && D > 0 && D % 8 == 0 && id_period > 0, "gather_rows: bad sizes rows=%d D=%d", rows, D);
  CGPT_REQUIRE(ids != nullptr || table_rows <= 0 || id_period <= table_rows,
               "gather_rows: id_period %d exceeds the table's %d rows", id_period, table_rows);
  CGPT_REQUIRE(ldt % 8 == 0 && ldo % 8 == 0, "gather_rows: leading dims must be multiples of 8");
  if (out_dtype == CGPT_DT_F32)
    gather_rows_kernel<true><<<rows, 128, 0, stream>>>((const __nv_bfloat16*)table, ldt, ids, id_period, D, out,
                                                       ldo, remap_period, remap_stride, remap_offset, table_rows);
  else
    gather_rows_kernel<false><<<rows, 128, 0, stream>>>((const __nv_bfloat16*)table, ldt, ids, id_period, D, out,
                                                        ldo, remap_period, remap_stride, remap_offset, table_rows);
  CGPT_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace cgpt
