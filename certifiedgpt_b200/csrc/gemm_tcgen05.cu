// Dense bf16 GEMM core for sm_100a:  out[M,N] = epilogue(A[M,K] . W[N,K]^T)
//
// Every nn.Linear / Conv2d-as-GEMM on the smoothing hot path goes through this kernel
// (SURVEY.md 2.3 rows K2,K3,K5,K6,K7,K9-K12,K14,K15; reference call sites
// eva_vit.py:126-131,151,60-64,202; Qformer.py:128-135,285-289,358-375; minigpt4.py:141;
// HF LlamaForCausalLM linears).  Both operands are K-major, which is the native
// nn.Linear weight layout ([out,in]), so no transposes exist anywhere.
//
// Structure (one persistent CTA per SM, 256 threads):
//   warp 8      TMA producer   : cp.async.bulk.tensor 128B-swizzled A/B k-blocks -> smem ring
//   warp 9      MMA issuer     : one thread issues tcgen05.mma (UMMA 128 x BN x 16), fp32
//                                accumulators in TMEM, two accumulator stages
//   warp 10     TMEM allocator
//   (k-blocks are 128 wide = two swizzle atoms, 8 MMAs per barrier round trip: with 64-wide blocks the
//    single issuing thread's wait/fence/commit chain left only ~13 % slack against the 512-cycle MMA
//    time of a block, and an ALU-heavy epilogue (GELU) sharing its scheduler pushed it past that:
//    14.0k instead of 11.3k cycles per 256x256x1408 tile, measured with CGPT_GEMM_DBG)
//   warps 0..7  epilogue       : 2 warps per TMEM lane quarter (column halves); software-pipelined
//                                tcgen05.ld -> bias (staged in smem) / GELU / SwiGLU / residual
//                                (prefetched) / pos-embed -> bf16 or fp32 global stores, overlapped
//                                with the next tile's MMAs through the second TMEM stage
#include <stdlib.h>
#include <deque>
#include "common.cuh"
#include "ops.h"

namespace cgpt {

constexpr int BM = 128;
constexpr int BK = 128;     // k-block per pipeline stage = KSUB swizzle atoms of 64 bf16
constexpr int BKA = 64;     // one 128-byte swizzle atom along K (TMA box width)
constexpr int KSUB = BK / BKA;
constexpr int UMMA_K = 16;
constexpr int GEMM_THREADS = 384;
constexpr int EPI_WARPS = 8;

// CTAS = 1: one CTA owns a 128 x BN tile.  CTAS = 2: a CTA pair (cluster of 2, cta_group::2) owns a
// 256 x BN tile; each CTA stages its own 128 rows of A and HALF of the B tile (BN/2 rows), so shared
// memory write (TMA) and read (UMMA) traffic per SM drops by a third and L2->SM traffic by a third.
// MT = 2 (CTA pairs only): each CTA owns TWO 128-row sub-tiles of A that share one B tile - a 512 x BN tile per pair, both
// TMEM accumulator stages belong to the same tile (no epilogue / mainloop overlap across tiles).  A third less L2 -> SM
// traffic and fewer DRAM re-reads per FLOP; k-blocks of 64 so that four 48 KB stages fit.  This is cuBLAS's own tiling
// for these shapes (nvjet 256x256 per CTA, 2cta; profiles/r02_gemm_vs_cublas.txt): the step runs at the 1 kW power cap,
// where bytes moved per FLOP count; the exposed epilogue is paid for in time, so the tiling is used only where the
// mainloop is long (K >= 8192, see the dispatch in gemm()).
template <int BN, int CTAS, int MT = 1>
struct GemmCfg {
  static constexpr int BK_ = MT == 2 ? 64 : BK;          // k-block of a pipeline stage
  static constexpr int KSUB_ = BK_ / BKA;
  static constexpr int A_SUB = BM * BKA * 2;             // one [128 x 64] sub-tile
  static constexpr int A_BYTES = A_SUB * KSUB_ * MT;
  static constexpr int B_ROWS = BN / CTAS;
  static constexpr int B_SUB = B_ROWS * BKA * 2;
  static constexpr int B_BYTES = B_SUB * KSUB_;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (200 * 1024) / STAGE_BYTES > 8 ? 8 : (200 * 1024) / STAGE_BYTES;
  static constexpr int RED_STAGE_BYTES = 8 * 32 * 20 * 4;   // 8 epilogue warps x [32 rows][16 + 4 pad] fp32
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + 2 * 256 * 4 /*bias*/ +
                                    RED_STAGE_BYTES + 2 * 32 * 8 /*head-major column offsets*/;
  static_assert(B_SUB % 1024 == 0, "B stage must keep 1024B alignment for SWIZZLE_128B");
  static_assert(BN % 16 == 0 && BN <= 256, "invalid UMMA N");
  static_assert(MT == 1 || (MT == 2 && CTAS == 2 && BN == 256), "two M sub-tiles per CTA: CTA pairs with 256-wide tiles only");
};

struct EpiParams {
  void* out;
  long long ldo;
  int out_f32;
  const float* bias;
  const void* resid;
  long long ldr;
  int resid_f32;
  int act;
  const float* row_add;
  long long ld_row_add;
  int row_period;     // >0: m -> (m / period, m % period)
  int row_add_offset; // row_add row = (m % period) + offset
  int remap_stride;   // out row = (m / period) * remap_stride + remap_offset + (m % period)
  int remap_offset;
  int red_inplace;    // out aliases the fp32 residual: accumulate with red.global.add (no residual load)
  int group_m;        // m-tiles per rasterisation band (tile_coords)
  // fused HF rotary embedding + KV-cache append for a fused QKV projection with 128-wide heads (ROPE kernels)
  int rope_T, rope_pos0, rope_hidden;          // rows per sample, position of row 0, heads * 128
  const float* rope_cos; const float* rope_sin; // [max_pos, 64]
  __nv_bfloat16* rope_k; __nv_bfloat16* rope_v; // KV cache
  long long rope_ldc; int rope_cache_rows, rope_cache_row0;
  // head-major scatter of a fused q|k|v projection (cgpt_gemm_epilogue.hm_*): out_row carries the (sample, token) part
  // of the address, the column part is worked out per 8-element piece (warp-uniform)
  int hm_T, hm_H, hm_hd, hm_D;                  // rows per sample, heads, head dim, H * hd
  float hm_inv_hd;
  long long hm_ws;                              // elements between the q, k and v blocks = M * D
  long long* dbg;     // optional per-CTA cycle counters (CGPT_GEMM_DBG): [grid][8]
};

// exact-erf GELU with erf from Abramowitz-Stegun 7.1.26 (|abs err| <= 1.5e-7, far below bf16
// resolution): 2 MUFU + ~12 FP32 ops instead of erff()'s ~28-instruction select chain
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float silu_fast(float x) {
  return x * rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * x));
}
__device__ __forceinline__ float gelu_erf_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = rcp_approx(fmaf(0.3275911f, z, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  p *= t;
  const float e = ex2_approx(-1.4426950408889634f * z * z);
  const float erf_abs = fmaf(-p, e, 1.0f);
  const float erf_x = copysignf(erf_abs, x);
  return 0.5f * x * (1.0f + erf_x);
}

// residual prefetch: 16 consecutive columns of one row into registers (fp32 values)
__device__ __forceinline__ void load_resid16(const EpiParams& p, long long out_row, int n0, float* r) {
  if (p.resid_f32) {
    const float* src = reinterpret_cast<const float*>(p.resid) + out_row * p.ldr + n0;
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
      const float4 b = *reinterpret_cast<const float4*>(src + i);
      r[i] = b.x; r[i + 1] = b.y; r[i + 2] = b.z; r[i + 3] = b.w;
    }
  } else {
    const __nv_bfloat16* src = reinterpret_cast<const __nv_bfloat16*>(p.resid) + out_row * p.ldr + n0;
#pragma unroll
    for (int i = 0; i < 16; i += 8) {
      const uint4 b = *reinterpret_cast<const uint4*>(src + i);
      r[i] = bf16_lo(b.x); r[i + 1] = bf16_hi(b.x); r[i + 2] = bf16_lo(b.y); r[i + 3] = bf16_hi(b.y);
      r[i + 4] = bf16_lo(b.z); r[i + 5] = bf16_hi(b.z); r[i + 6] = bf16_lo(b.w); r[i + 7] = bf16_hi(b.w);
    }
  }
}

// process 16 consecutive accumulator columns of one row and store them (HM: head-major q|k|v scatter, its own kernel
// instantiation so that the hot epilogue loop of every other GEMM keeps its code size)
template <bool HM>
__device__ __forceinline__ void epilogue_store16(const uint32_t* acc, const EpiParams& p, const float* bias_s,
                                                 const float* resid_r, int m, long long out_row, int n0,
                                                 const long long* hm_off = nullptr) {
  constexpr int NC = 16;
  float v[NC];
#pragma unroll
  for (int i = 0; i < NC; ++i) v[i] = __uint_as_float(acc[i]);
  if (p.bias != nullptr) {
#pragma unroll
    for (int i = 0; i < NC; i += 4) {
      const float4 b = *reinterpret_cast<const float4*>(bias_s + i);   // smem broadcast
      v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
    }
  }
  if (p.act == CGPT_ACT_GELU) {
#pragma unroll
    for (int i = 0; i < NC; ++i) v[i] = gelu_erf_fast(v[i]);
  }
  if (p.act == CGPT_ACT_QUICKGELU) {
    // x * sigmoid(1.702 x) (CLIP's quick_gelu)
#pragma unroll
    for (int i = 0; i < NC; ++i) v[i] = v[i] * rcp_approx(1.0f + ex2_approx(-1.702f * 1.4426950408889634f * v[i]));
  }
  if (p.act == CGPT_ACT_SWIGLU) {
    // weight rows are interleaved (gate_j, up_j): out[:, j] = silu(gate_j) * up_j
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + out_row * p.ldo + (n0 >> 1);
    uint32_t w[NC / 4];
#pragma unroll
    for (int i = 0; i < NC / 4; ++i)
      w[i] = pack_bf16x2(silu_fast(v[4 * i]) * v[4 * i + 1], silu_fast(v[4 * i + 2]) * v[4 * i + 3]);
    __stcs(reinterpret_cast<uint4*>(o), make_uint4(w[0], w[1], w[2], w[3]));
    return;
  }
  if (p.red_inplace) {
    // in-place fp32 residual update (x += proj(...)): every element is touched by exactly one thread, so an
    // uncontended fire-and-forget L2 reduction replaces the latency-bound load + add + store (same single
    // fp32 addition, bit-identical result)
    float* o = reinterpret_cast<float*>(p.out) + out_row * p.ldo + n0;
#pragma unroll
    for (int i = 0; i < NC; i += 4)
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + i), "f"(v[i]), "f"(v[i + 1]),
                   "f"(v[i + 2]), "f"(v[i + 3])
                   : "memory");
    return;
  }
  if (p.row_add != nullptr) {
    const float* ra = p.row_add + (long long)((m % p.row_period) + p.row_add_offset) * p.ld_row_add + n0;
#pragma unroll
    for (int i = 0; i < NC; i += 4) {
      float4 b = __ldg(reinterpret_cast<const float4*>(ra + i));
      v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
    }
  }
  if (p.resid != nullptr) {
#pragma unroll
    for (int i = 0; i < NC; ++i) v[i] += resid_r[i];
  }
  if constexpr (HM) {
    // hm_off[i]: offset of the i-th 8-column piece of this chunk (shared-memory table, one entry per piece and tile)
#pragma unroll
    for (int i = 0; i < NC; i += 8) {
      __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + out_row + hm_off[i >> 3];
      __stcs(reinterpret_cast<uint4*>(o),
             make_uint4(pack_bf16x2(v[i], v[i + 1]), pack_bf16x2(v[i + 2], v[i + 3]),
                        pack_bf16x2(v[i + 4], v[i + 5]), pack_bf16x2(v[i + 6], v[i + 7])));
    }
    return;
  }
  if (p.out_f32) {
    float* o = reinterpret_cast<float*>(p.out) + out_row * p.ldo + n0;
#pragma unroll
    for (int i = 0; i < NC; i += 4)
      __stcs(reinterpret_cast<float4*>(o + i), make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]));
  } else {
    // streaming (evict-first) stores: the output is consumed by the NEXT kernel after GBs of other traffic,
    // so it should not push this GEMM's re-used A band out of L2
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + out_row * p.ldo + n0;
#pragma unroll
    for (int i = 0; i < NC; i += 8)
      __stcs(reinterpret_cast<uint4*>(o + i),
             make_uint4(pack_bf16x2(v[i], v[i + 1]), pack_bf16x2(v[i + 2], v[i + 3]),
                        pack_bf16x2(v[i + 4], v[i + 5]), pack_bf16x2(v[i + 6], v[i + 7])));
  }
}

// one epilogue warp's share of a tile: chunks [c0, c1) of 16 columns.  The TMEM load of the next
// chunk and its residual prefetch are in flight while the current chunk is processed; the loop is
// unrolled by exactly 2 (static register double buffer) and NOT further, so the SASS stays small
// enough for the instruction cache with 8 warps running it.
template <bool HM>
__device__ __forceinline__ void epilogue_chunks(uint32_t t_row, const EpiParams& epi, const float* bias_s,
                                                int m, long long out_row, bool row_ok, int n_base, int N,
                                                int c0, int c1, const long long* hm_tbl = nullptr) {
  uint32_t r0[16], r1[16];
  float rr0[16], rr1[16];
  const bool has_res = epi.resid != nullptr && row_ok && !epi.red_inplace;
  tmem_ld_x16(t_row + c0 * 16, r0);
  if (has_res && n_base + c0 * 16 < N) load_resid16(epi, out_row, n_base + c0 * 16, rr0);
#pragma unroll 1
  for (int c = c0; c < c1; c += 2) {
    tmem_ld_wait();
    if (c + 1 < c1) {
      tmem_ld_x16(t_row + (c + 1) * 16, r1);
      if (has_res && n_base + (c + 1) * 16 < N) load_resid16(epi, out_row, n_base + (c + 1) * 16, rr1);
    }
    if (row_ok && n_base + c * 16 < N)
      epilogue_store16<HM>(r0, epi, bias_s + c * 16, rr0, m, out_row, n_base + c * 16, HM ? hm_tbl + 2 * c : nullptr);
    if (c + 1 < c1) {
      tmem_ld_wait();
      if (c + 2 < c1) {
        tmem_ld_x16(t_row + (c + 2) * 16, r0);
        if (has_res && n_base + (c + 2) * 16 < N) load_resid16(epi, out_row, n_base + (c + 2) * 16, rr0);
      }
      if (row_ok && n_base + (c + 1) * 16 < N)
        epilogue_store16<HM>(r1, epi, bias_s + (c + 1) * 16, rr1, m, out_row, n_base + (c + 1) * 16,
                             HM ? hm_tbl + 2 * (c + 1) : nullptr);
    }
  }
}

// In-place fp32 residual update (x += proj(...)) with COALESCED L2 reductions.  The TMEM layout gives every thread one
// row, so a direct red.global.add.v4.f32 per thread touches 32 different rows per warp instruction: 32 half-used
// sectors, and the L2 reduction path saturates (the short-K ViT proj GEMM spent 14k cycles per tile here against 11k
// of MMA).  Instead each warp transposes its [32 rows x 16 cols] chunk through a private shared-memory patch so that
// 4 adjacent lanes cover 64 contiguous bytes of one row: 8 rows x 2 full sectors per instruction, 4x fewer L2
// requests.  Same single fp32 addition per element as before (every element is still touched by exactly one thread).
// RMW = true replaces the reduction by a plain load + add + store in that transposed domain: the per-SM issue rate of
// RED (~1.3 cycles per lane and request, B300_MICROARCH "Atomics") made 8192 requests per 128 x 256 tile cost ~10.6k
// cycles against 11.3k cycles of MMA for K = 1408, while 256 coalesced 16-byte loads + 256 stores cost ~2k.
template <bool RMW>
__device__ __forceinline__ void epilogue_chunks_red(uint32_t t_row, const EpiParams& p, const float* bias_s, float* stage,
                                                    int m_warp0, int M, int n_base, int N, int c0, int c1) {
  const int lane = threadIdx.x & 31;
  uint32_t r0[16], r1[16];
  const int srow = lane >> 2, scol = (lane & 3) * 4;        // this lane's (row within a group of 8, column) on the way out
  // the 4 output rows this lane serves in every chunk (k * 8 + srow), fixed for the whole tile
  float* orow[4];
  bool rok[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int m = m_warp0 + k * 8 + srow;
    rok[k] = m < M;
    long long o = m;
    if (p.row_period > 0 && p.remap_stride > 0)
      o = (long long)(m / p.row_period) * p.remap_stride + p.remap_offset + (m % p.row_period);
    orow[k] = reinterpret_cast<float*>(p.out) + o * p.ldo + scol;
  }
  tmem_ld_x16(t_row + c0 * 16, r0);
#pragma unroll 1
  for (int c = c0; c < c1; ++c) {
    const int n0 = n_base + c * 16;
    // RMW form: the residual values of this chunk are requested BEFORE the TMEM wait and the shared-memory transpose,
    // so most of their L2 latency hides behind those steps (4 independent 16-byte loads per lane, 64 contiguous bytes
    // per row and instruction)
    float4 old[4];
    if (RMW && n0 < N) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (rok[k]) old[k] = *reinterpret_cast<const float4*>(orow[k] + n0);
    }
    tmem_ld_wait();
    uint32_t* cur = ((c - c0) & 1) ? r1 : r0;
    uint32_t* nxt = ((c - c0) & 1) ? r0 : r1;
    if (c + 1 < c1) tmem_ld_x16(t_row + (c + 1) * 16, nxt);
    if (n0 < N) {                                           // warp-uniform (N is a multiple of 16)
      float* mine = stage + lane * 20;
#pragma unroll
      for (int i = 0; i < 16; i += 4) {
        float4 v = make_float4(__uint_as_float(cur[i]), __uint_as_float(cur[i + 1]), __uint_as_float(cur[i + 2]),
                               __uint_as_float(cur[i + 3]));
        if (p.bias != nullptr) {
          const float4 b = *reinterpret_cast<const float4*>(bias_s + c * 16 + i);
          v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
        }
        *reinterpret_cast<float4*>(mine + i) = v;
      }
      __syncwarp();
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (rok[k]) {
          const float4 v = *reinterpret_cast<const float4*>(stage + (k * 8 + srow) * 20 + scol);
          float* o = orow[k] + n0;
          if (RMW) {
            // same single fp32 addition per element as the reduction form (resid + (acc + bias)): bit-identical
            *reinterpret_cast<float4*>(o) = make_float4(old[k].x + v.x, old[k].y + v.y, old[k].z + v.z, old[k].w + v.w);
          } else {
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                         : "memory");
          }
        }
      }
      __syncwarp();
    }
  }
}

// Fused QKV epilogue (ROPE kernels): this warp's 128 accumulator columns are exactly one head of q, k or v.
// q and k get HF's rotary embedding (rotate_half convention: out[j] = x[j] cos_j - x[j+64] sin_j,
// out[j+64] = x[j+64] cos_j + x[j] sin_j) on the fp32 accumulators; q goes to the q part of `out`, rotated k and
// plain v go straight to the KV cache row of (sample, position): the separate rope / cache-append pass and its
// round trip through the QKV buffer disappear.
__device__ __forceinline__ void epilogue_rope_head(uint32_t t_row, const EpiParams& p, int m, bool row_ok, int col0) {
  const int region = col0 / p.rope_hidden;            // 0 = q, 1 = k, 2 = v
  const int cin = col0 - region * p.rope_hidden;      // column inside q / k / v
  const int b = m / p.rope_T, i = m - b * p.rope_T;
  const int pos = p.rope_pos0 + i;
  __nv_bfloat16* dst;
  if (region == 0) {
    dst = reinterpret_cast<__nv_bfloat16*>(p.out) + static_cast<long long>(m) * p.ldo + cin;
  } else {
    const long long crow = (static_cast<long long>(b) * p.rope_cache_rows + p.rope_cache_row0 + i) * p.rope_ldc;
    dst = (region == 1 ? p.rope_k : p.rope_v) + crow + cin;
  }
  const float* cs = p.rope_cos + static_cast<long long>(pos) * 64;
  const float* sn = p.rope_sin + static_cast<long long>(pos) * 64;
  uint32_t lo[16], hi[16];
#pragma unroll 1
  for (int c = 0; c < 4; ++c) {
    tmem_ld_x16(t_row + c * 16, lo);
    tmem_ld_x16(t_row + 64 + c * 16, hi);
    tmem_ld_wait();
    if (!row_ok) continue;
    uint32_t w1[8], w2[8];
    if (region < 2) {
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        const float4 c4 = __ldg(reinterpret_cast<const float4*>(cs + c * 16 + j));
        const float4 s4 = __ldg(reinterpret_cast<const float4*>(sn + c * 16 + j));
        const float cc[4] = {c4.x, c4.y, c4.z, c4.w}, ss[4] = {s4.x, s4.y, s4.z, s4.w};
        float o1[4], o2[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float x1 = __uint_as_float(lo[j + t]), x2 = __uint_as_float(hi[j + t]);
          o1[t] = x1 * cc[t] - x2 * ss[t];
          o2[t] = x2 * cc[t] + x1 * ss[t];
        }
        w1[j / 2] = pack_bf16x2(o1[0], o1[1]); w1[j / 2 + 1] = pack_bf16x2(o1[2], o1[3]);
        w2[j / 2] = pack_bf16x2(o2[0], o2[1]); w2[j / 2 + 1] = pack_bf16x2(o2[2], o2[3]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; j += 2) {
        w1[j / 2] = pack_bf16x2(__uint_as_float(lo[j]), __uint_as_float(lo[j + 1]));
        w2[j / 2] = pack_bf16x2(__uint_as_float(hi[j]), __uint_as_float(hi[j + 1]));
      }
    }
    uint4* d1 = reinterpret_cast<uint4*>(dst + c * 16);
    uint4* d2 = reinterpret_cast<uint4*>(dst + 64 + c * 16);
    d1[0] = make_uint4(w1[0], w1[1], w1[2], w1[3]);
    d1[1] = make_uint4(w1[4], w1[5], w1[6], w1[7]);
    d2[0] = make_uint4(w2[0], w2[1], w2[2], w2[3]);
    d2[1] = make_uint4(w2[4], w2[5], w2[6], w2[7]);
  }
}

// Tile rasterisation: bands of group_m m-tiles, m fastest inside a band.  The CTAs running at any
// moment then share ~group_m A tiles and only ~(#CTAs / group_m) weight tiles, so the working set
// stays inside the 126 MB L2 even for the 180 MB Llama gate/up weight (n-fastest order re-read the
// whole weight from HBM once per m-tile: 35 GB of DRAM reads for 0.77 GB of operands, ncu r01).
constexpr int GROUP_M_MAX = 16;
__device__ __forceinline__ void tile_coords(int tile, int m_tiles, int n_tiles, int group_m, int& m, int& n) {
  const int band = group_m * n_tiles;
  const int g = tile / band;
  const int first_m = g * group_m;
  const int gm = min(group_m, m_tiles - first_m);
  const int r = tile - g * band;
  m = first_m + r % gm;
  n = r / gm;
}

// MODE: 0 = the common epilogues, 1 = fused rotary + KV-cache append (ROPE), 2 = head-major q|k|v scatter
//       3 = in-place fp32 residual update as a coalesced load + add + store (the short-K ViT proj GEMM)
// Each special epilogue is its own instantiation so that the hot epilogue loop of the common kernel keeps the code size
// it was tuned at (the GELU GEMM lost 19 % when two more branches were compiled into it: instruction cache).
constexpr int MODE_PLAIN = 0, MODE_ROPE = 1, MODE_HM = 2, MODE_RMW = 3;
template <int BN, int CTAS, int MODE = MODE_PLAIN, int MT = 1>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tma_a,
                         const __grid_constant__ CUtensorMap tma_b, int M, int N, int K,
                         EpiParams epi) {
  using Cfg = GemmCfg<BN, CTAS, MT>;
  constexpr int BKC = Cfg::BK_, KSUBC = Cfg::KSUB_;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by pointer arithmetic on the __shared__ array: a round trip through uintptr_t would
  // lose the address space and turn every shared-memory access below into a generic LD.E / ST.E
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * Cfg::A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* full_bar = bars;                  // [STAGES] TMA -> MMA
  uint64_t* empty_bar = bars + STAGES;        // [STAGES] MMA -> TMA
  uint64_t* tmem_full = bars + 2 * STAGES;    // [2] MMA -> epilogue
  uint64_t* tmem_empty = bars + 2 * STAGES + 2;  // [2] epilogue -> MMA
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  float* bias_smem = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES + 256);  // [2][256]
  float* red_stage = bias_smem + 2 * 256;                                                // [8 warps][32][20]
  long long* hm_tbl = reinterpret_cast<long long*>(red_stage + 8 * 32 * 20);             // [2 acc stages][32 pieces of 8 columns]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // a "tile" is (BM*CTAS) x BN; with CTAS = 2 both CTAs of the pair walk the same tile sequence
  const int cta_rank = CTAS == 2 ? static_cast<int>(cluster_ctarank()) : 0;
  const int worker = blockIdx.x / CTAS, num_workers = gridDim.x / CTAS;
  const int m_tiles = (M + BM * CTAS * MT - 1) / (BM * CTAS * MT);
  const int n_tiles = (N + BN - 1) / BN;
  const int num_tiles = m_tiles * n_tiles;
  const int num_kb = (K + BKC - 1) / BKC;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
  }
  if (warp == 9 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], EPI_WARPS * CTAS);   // both CTAs' epilogues release the leader's MMA
    }
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 10) {
    if (CTAS == 2) { tmem_alloc_2sm(tmem_ptr, 512); tmem_relinquish_2sm(); }
    else           { tmem_alloc(tmem_ptr, 512); tmem_relinquish(); }
  }
  tcgen05_fence_before();
  __syncwarp();
  if (CTAS == 2) { __syncthreads(); cluster_sync_all(); } else __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 8) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = worker; tile < num_tiles; tile += num_workers) {
        int m_blk, n_blk;
        tile_coords(tile, m_tiles, n_tiles, epi.group_m, m_blk, n_blk);
        // 128-row block of sub-tile mt of this CTA: ((m_tile * MT + mt) * CTAS + cta_rank)
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (CTAS == 2) {
            // the leader's barrier collects the bytes of BOTH CTAs' loads
            if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
#pragma unroll
            for (int sub = 0; sub < KSUBC; ++sub) {
#pragma unroll
              for (int mt = 0; mt < MT; ++mt)
                tma_load_2d_2sm(smem_a + stage * Cfg::A_BYTES + (mt * KSUBC + sub) * Cfg::A_SUB, &tma_a, &full_bar[stage],
                                kb * BKC + sub * BKA, ((m_blk * MT + mt) * CTAS + cta_rank) * BM);
              tma_load_2d_2sm(smem_b + stage * Cfg::B_BYTES + sub * Cfg::B_SUB, &tma_b, &full_bar[stage],
                              kb * BKC + sub * BKA, n_blk * BN + cta_rank * Cfg::B_ROWS);
            }
          } else {
            mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
#pragma unroll
            for (int sub = 0; sub < KSUBC; ++sub) {
              tma_load_2d(smem_a + stage * Cfg::A_BYTES + sub * Cfg::A_SUB, &tma_a, &full_bar[stage],
                          kb * BKC + sub * BKA, (m_blk * CTAS + cta_rank) * BM);
              tma_load_2d(smem_b + stage * Cfg::B_BYTES + sub * Cfg::B_SUB, &tma_b, &full_bar[stage],
                          kb * BKC + sub * BKA, n_blk * BN);
            }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 9) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0 && cta_rank == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM * CTAS, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      long long t_wait_epi = 0, t_wait_tma = 0, t_total0 = clock64();
      for (int tile = worker; tile < num_tiles; tile += num_workers) {
        long long c0 = clock64();
        if (MT == 2) { mbar_wait(&tmem_empty[0], acc_phase ^ 1); mbar_wait(&tmem_empty[1], acc_phase ^ 1); }
        else mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        t_wait_epi += clock64() - c0;
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + (MT == 2 ? 0 : acc * 256);
        for (int kb = 0; kb < num_kb; ++kb) {
          c0 = clock64();
          mbar_wait(&full_bar[stage], phase);
          t_wait_tma += clock64() - c0;
          tcgen05_fence_after();
          const uint64_t adesc = make_smem_desc_sw128(smem_u32(smem_a + stage * Cfg::A_BYTES));
          const uint64_t bdesc = make_smem_desc_sw128(smem_u32(smem_b + stage * Cfg::B_BYTES));
#pragma unroll
          for (int k = 0; k < BKC / UMMA_K; ++k) {
            // sub-tile (k / 4) of the stage, then 16 bf16 = 32 B inside the 128B swizzle atom (+2 in addr>>4 units)
            const uint64_t bd = bdesc + (k >> 2) * (Cfg::B_SUB >> 4) + 2 * (k & 3);
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {      // the M sub-tiles of the tile share the B operand; accumulator mt
              const uint64_t ad = adesc + (mt * KSUBC + (k >> 2)) * (Cfg::A_SUB >> 4) + 2 * (k & 3);
              if (CTAS == 2) umma_bf16_2sm(d_tmem + mt * 256, ad, bd, idesc, (kb | k) != 0);
              else           umma_bf16(d_tmem, ad, bd, idesc, (kb | k) != 0);
            }
          }
          if (CTAS == 2) umma_commit_2sm(&empty_bar[stage]); else umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (MT == 2) {
          umma_commit_2sm(&tmem_full[0]);
          umma_commit_2sm(&tmem_full[1]);
          acc_phase ^= 1;
        } else {
          if (CTAS == 2) umma_commit_2sm(&tmem_full[acc]); else umma_commit(&tmem_full[acc]);
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
      }
      if (epi.dbg) {
        epi.dbg[blockIdx.x * 8 + 0] = t_wait_epi;
        epi.dbg[blockIdx.x * 8 + 1] = t_wait_tma;
        epi.dbg[blockIdx.x * 8 + 2] = clock64() - t_total0;
      }
    }
  } else if (warp < 8) {
    // ------------------------------------------------------------ epilogue (8 warps)
    const int quarter = warp & 3;          // TMEM lanes [32*quarter, 32*quarter+32)
    const int half = warp >> 2;            // column half of the tile
    const int epi_tid = threadIdx.x;       // 0..255
    constexpr int NCH = BN / 16;
    constexpr int NCH0 = (NCH + 1) / 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    long long t_wait_mma = 0, t_work = 0, t_bar = 0;
    for (int tile = worker; tile < num_tiles; tile += num_workers)
    for (int mt = 0; mt < MT; ++mt) {          // MT = 2: accumulator stage mt holds the tile's M sub-tile mt
      int m_blk, n_blk;
      tile_coords(tile, m_tiles, n_tiles, epi.group_m, m_blk, n_blk);
      m_blk = (m_blk * MT + mt) * CTAS + cta_rank;
      const int n_base = n_blk * BN;
      float* bias_s = bias_smem + acc * 256;
      long long c0 = clock64();
      if (epi.bias != nullptr && epi_tid < BN) {
        const int n = n_base + epi_tid;
        bias_s[epi_tid] = n < N ? __ldg(epi.bias + n) : 0.f;
      }
      if constexpr (MODE == MODE_HM) {
        // element offset of every 8-column piece of this tile inside the head-major output, minus the row part:
        // worked out once per tile by 32 threads instead of once per piece and thread (8-element pieces never straddle a
        // head: hd % 8 == 0)
        if (epi_tid < BN / 8) {
          const int n = n_base + epi_tid * 8;
          const int which = (n >= epi.hm_D) + (n >= 2 * epi.hm_D);
          const int rem = n - which * epi.hm_D;
          const int h = static_cast<int>((static_cast<float>(rem) + 0.5f) * epi.hm_inv_hd);
          hm_tbl[acc * 32 + epi_tid] = which * epi.hm_ws + static_cast<long long>(h) * epi.hm_T * epi.hm_hd + (rem - h * epi.hm_hd);
        }
      }
      {
        // pull this thread's residual row segment towards L2 while the tile's MMAs still run: the
        // row-per-thread residual loads below are latency-bound (25k cycles per tile vs 11k of MMA otherwise)
        const int mp = m_blk * BM + quarter * 32 + lane;
        if (epi.resid != nullptr && !epi.red_inplace && mp < M) {
          long long orow = mp;
          if (epi.row_period > 0 && epi.remap_stride > 0)
            orow = (long long)(mp / epi.row_period) * epi.remap_stride + epi.remap_offset + (mp % epi.row_period);
          const int c_lo = half == 0 ? 0 : NCH0 * 16, c_hi = half == 0 ? NCH0 * 16 : BN;
          const int esz = epi.resid_f32 ? 4 : 2;
          const char* rp = reinterpret_cast<const char*>(epi.resid) + (orow * epi.ldr + n_base) * esz;
          for (int c = c_lo; c < c_hi && n_base + c < N; c += 128 / esz)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(rp + (long long)c * esz));
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");   // epilogue warps only
      long long c1 = clock64();
      t_bar += c1 - c0;
      mbar_wait(&tmem_full[acc], acc_phase);
      c0 = clock64();
      t_wait_mma += c0 - c1;
      tcgen05_fence_after();
      const int m = m_blk * BM + quarter * 32 + lane;
      const bool row_ok = m < M;
      long long out_row = m;
      if (epi.row_period > 0 && epi.remap_stride > 0)
        out_row = (long long)(m / epi.row_period) * epi.remap_stride + epi.remap_offset +
                  (m % epi.row_period);
      if constexpr (MODE == MODE_HM) {
        // head-major: element offset of (sample b, head 0, token t, d 0) inside one of the q / k / v blocks
        const int b = m / epi.hm_T, t = m - b * epi.hm_T;
        out_row = (static_cast<long long>(b) * epi.hm_H * epi.hm_T + t) * epi.hm_hd;
      }
      const uint32_t t_row = tmem_base + acc * 256 + (static_cast<uint32_t>(quarter * 32) << 16);
      if constexpr (MODE == MODE_ROPE) {
        // BN = 256: this warp's half of the tile is one 128-wide head
        if (n_base + half * 128 < N) epilogue_rope_head(t_row + half * 128, epi, m, row_ok, n_base + half * 128);
      } else if constexpr (MODE == MODE_RMW) {
        epilogue_chunks_red<true>(t_row, epi, bias_s, red_stage + warp * (32 * 20), m_blk * BM + quarter * 32, M, n_base, N,
                                  half == 0 ? 0 : NCH0, half == 0 ? NCH0 : NCH);
      } else if (epi.red_inplace == 2) {
        epilogue_chunks_red<false>(t_row, epi, bias_s, red_stage + warp * (32 * 20), m_blk * BM + quarter * 32, M, n_base, N,
                                   half == 0 ? 0 : NCH0, half == 0 ? NCH0 : NCH);
      } else {
        epilogue_chunks<MODE == MODE_HM>(t_row, epi, bias_s, m, out_row, row_ok, n_base, N, half == 0 ? 0 : NCH0,
                                         half == 0 ? NCH0 : NCH, hm_tbl + acc * 32);
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) { if (CTAS == 2) mbar_arrive_leader(&tmem_empty[acc]); else mbar_arrive(&tmem_empty[acc]); }
      t_work += clock64() - c0;
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (epi.dbg && threadIdx.x == 0) {
      epi.dbg[blockIdx.x * 8 + 4] = t_wait_mma;
      epi.dbg[blockIdx.x * 8 + 5] = t_work;
      epi.dbg[blockIdx.x * 8 + 6] = t_bar;
    }
  }

  tcgen05_fence_before();
  __syncwarp();
  __syncthreads();
  if (CTAS == 2) cluster_sync_all();   // peer smem / barriers stay valid until both CTAs are done
  if (warp == 10) {
    tcgen05_fence_after();
    if (CTAS == 2) tmem_dealloc_2sm(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

// K-major bf16 matrix [rows, K] with leading dimension ld (elements); box = 64 x box_rows, SW128
static int make_tmap(CUtensorMap* map, const void* ptr, long long rows, long long K, long long ld,
                     int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  CGPT_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled driver entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)BKA, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CGPT_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d): rows=%lld K=%lld ld=%lld ptr=%p",
               (int)r, rows, K, ld, ptr);
  return 0;
}

static int g_num_sms = 0;
static int g_gemm_launches = 0;

struct GemmProfRec { cudaEvent_t e0, e1; int M, N, K; };
static std::deque<GemmProfRec> g_prof;
static bool g_prof_on = false;

template <int BN, int CTAS, int MODE = MODE_PLAIN, int MT = 1>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, int M, int N, int K,
                       const EpiParams& epi, int max_ctas, cudaStream_t stream) {
  using Cfg = GemmCfg<BN, CTAS, MT>;
  static bool configured = false;
  if (!configured) {
    CGPT_CHECK_CUDA(cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<BN, CTAS, MODE, MT>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    configured = true;
  }
  const int tiles = ((M + BM * CTAS * MT - 1) / (BM * CTAS * MT)) * ((N + BN - 1) / BN);
  int workers = g_num_sms / CTAS;
  if (max_ctas > 0 && workers > max_ctas / CTAS) workers = max_ctas / CTAS > 0 ? max_ctas / CTAS : 1;
  if (tiles < workers) workers = tiles;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(workers * CTAS);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CTAS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CGPT_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_bf16_tcgen05_kernel<BN, CTAS, MODE, MT>, ta, tb, M, N, K, epi));
  ++g_gemm_launches;
  return 0;
}

int gemm_launch_count() { return g_gemm_launches; }

// pick the N tile: exact divisors first (no wasted MMA columns), else 256 with masking
static int pick_bn(int N) {
  if (N % 256 == 0) return 256;
  if (N % 176 == 0) return 176;
  if (N % 128 == 0) return 128;
  return N >= 256 ? 256 : (N > 128 ? 176 : 128);
}

// Skinny problems (one or two m-tiles: the decode steps, and every GEMM of a small per-rank batch) are bound by how
// many SMs stream the weight, not by the tensor pipe: with 256-wide tiles an N = 4096 projection keeps 16 CTA pairs
// busy and N = 22016 runs 74 + 12 tiles.  Pick the N tile that minimises waves x (tile width + fixed cost), and leave
// 256 unless the model predicts >= 15 % (measured at M = 138: down 71.7 -> 53.2 us with 128-wide tiles,
// scripts/gemm_wave_probe.py).  The tile width does not change any element's K summation order, so results stay
// bit-identical across batch sizes.  CGPT_GEMM_NO_SKINNY=1 disables.
static int pick_bn_skinny(int M, int N, int ctas, int bn_default, int workers) {
  static const bool off = getenv("CGPT_GEMM_NO_SKINNY") != nullptr;
  const int m_tiles = (M + BM * ctas - 1) / (BM * ctas);
  if (off || m_tiles > 2 || N < 512 || workers <= 0) return bn_default;
  auto cost = [&](int bn) {
    const long long tiles = static_cast<long long>(m_tiles) * ((N + bn - 1) / bn);
    return ((tiles + workers - 1) / workers) * (bn + 64);
  };
  const long long cost_default = cost(bn_default);
  long long best_cost = cost_default;
  int best = bn_default;
  const int cand[3] = {256, 176, 128};
  for (int i = 0; i < 3; ++i) {
    const long long c = cost(cand[i]);
    if (c < best_cost) { best_cost = c; best = cand[i]; }
  }
  return best_cost * 100 <= cost_default * 85 ? best : bn_default;
}

int gemm_bf16(const void* A, long long lda, const void* W, long long ldw, int M, int N, int K,
              const cgpt_gemm_epilogue* e, int force_bn, cudaStream_t stream) {
  CGPT_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: empty problem M=%d N=%d K=%d", M, N, K);
  CGPT_REQUIRE(K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0,
               "gemm: K, lda, ldw must be multiples of 8 (16-byte TMA rows): K=%d lda=%lld ldw=%lld",
               K, lda, ldw);
  CGPT_REQUIRE(N % 16 == 0, "gemm: N must be a multiple of 16 (got %d)", N);
  CGPT_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0,
               "gemm: A and W must be 16-byte aligned");
  CGPT_REQUIRE(e != nullptr && e->out != nullptr, "gemm: epilogue/out is null");
  if (g_num_sms == 0) {
    int dev = 0;
    CGPT_CHECK_CUDA(cudaGetDevice(&dev));
    CGPT_CHECK_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  EpiParams p;
  p.out = e->out; p.ldo = e->ldo; p.out_f32 = e->out_dtype == CGPT_DT_F32;
  p.bias = e->bias;
  p.resid = e->resid; p.ldr = e->ldr; p.resid_f32 = e->resid_dtype == CGPT_DT_F32;
  p.act = e->act;
  p.row_add = e->row_add; p.ld_row_add = e->ld_row_add;
  p.row_period = e->row_period; p.row_add_offset = e->row_add_offset;
  p.remap_stride = e->remap_stride; p.remap_offset = e->remap_offset;
  p.red_inplace = (p.resid != nullptr && p.resid == p.out && p.out_f32 && p.resid_f32 && p.ldr == p.ldo &&
                   p.act == CGPT_ACT_NONE && p.row_add == nullptr && !getenv("CGPT_GEMM_NO_RED")) ? 1 : 0;
  // 2 = warp-transposed, sector-coalesced reductions, for the SHORT-K GEMMs whose epilogue outlasts their MMAs (ViT
  // proj, K = 1408: 775 -> 1102 TFLOP/s).  Long-K GEMMs hide the direct form behind the MMAs, and inside the
  // power-capped step the extra shared-memory round trip cost them 1-4 % (A/B, gpurun_out/ab_red*.log), so they keep
  // it.  CGPT_GEMM_RED_DIRECT=1 / CGPT_GEMM_RED_COALESCED=1 force one form for experiments.
  if (p.red_inplace && p.ldo % 4 == 0 && !getenv("CGPT_GEMM_RED_DIRECT") &&
      (K <= 2048 || getenv("CGPT_GEMM_RED_COALESCED")))
    p.red_inplace = getenv("CGPT_GEMM_RED_RMW") ? 3 : 2;   // 3 = coalesced load + add + store instead of RED: measured SLOWER
                                                           // on the ViT proj GEMM (834 vs 952 TFLOP/s at batch 1024, one box,
                                                           // gpurun_out/r2_c4_prof_*.log): the loads sit on the tile's critical
                                                           // path, the reductions are fire-and-forget; kept as an experiment
  p.dbg = reinterpret_cast<long long*>(getenv("CGPT_GEMM_DBG") ? strtoull(getenv("CGPT_GEMM_DBG"), nullptr, 0) : 0ull);
  p.hm_T = 0; p.hm_H = 0; p.hm_hd = 0; p.hm_D = 0; p.hm_inv_hd = 0.f; p.hm_ws = 0;
  if (e->hm_T > 0) {
    CGPT_REQUIRE(e->hm_heads > 0 && e->hm_hd > 0 && e->hm_hd % 8 == 0 && N == 3 * e->hm_heads * e->hm_hd && M % e->hm_T == 0,
                 "gemm(head-major): N = %d must be 3 * heads (%d) * hd (%d, multiple of 8) and M = %d a multiple of T = %d", N,
                 e->hm_heads, e->hm_hd, M, e->hm_T);
    CGPT_REQUIRE(!p.out_f32 && p.resid == nullptr && p.act == CGPT_ACT_NONE && p.row_add == nullptr && p.remap_stride == 0 &&
                     e->rope == nullptr,
                 "gemm(head-major): the scatter epilogue takes bf16 out, optional bias, and nothing else");
    p.hm_T = e->hm_T; p.hm_H = e->hm_heads; p.hm_hd = e->hm_hd; p.hm_D = e->hm_heads * e->hm_hd;
    p.hm_inv_hd = 1.0f / static_cast<float>(e->hm_hd);
    p.hm_ws = static_cast<long long>(M) * p.hm_D;
  }
  const cgpt_gemm_rope* rp = e->rope;
  if (rp != nullptr) {
    CGPT_REQUIRE(rp->heads > 0 && N == 3 * rp->heads * 128, "gemm(rope): N = %d must be 3 * heads * 128 (heads = %d)", N, rp->heads);
    CGPT_REQUIRE(!p.out_f32 && p.bias == nullptr && p.resid == nullptr && p.act == CGPT_ACT_NONE && p.row_add == nullptr &&
                     p.remap_stride == 0,
                 "gemm(rope): the fused QKV epilogue takes bf16 out and no bias / residual / activation / remap");
    CGPT_REQUIRE(rp->T > 0 && M % rp->T == 0 && rp->cos_table && rp->sin_table && rp->kcache && rp->vcache &&
                     rp->ld_cache % 8 == 0 && p.ldo % 8 == 0 && rp->cache_row0 + rp->T <= rp->cache_rows_per_batch,
                 "gemm(rope): bad rope / cache arguments (M = %d, T = %d)", M, rp->T);
    p.rope_T = rp->T; p.rope_pos0 = rp->pos0; p.rope_hidden = rp->heads * 128;
    p.rope_cos = rp->cos_table; p.rope_sin = rp->sin_table;
    p.rope_k = reinterpret_cast<__nv_bfloat16*>(rp->kcache); p.rope_v = reinterpret_cast<__nv_bfloat16*>(rp->vcache);
    p.rope_ldc = rp->ld_cache; p.rope_cache_rows = rp->cache_rows_per_batch; p.rope_cache_row0 = rp->cache_row0;
  }
  CGPT_REQUIRE(p.row_add == nullptr || p.row_period > 0, "gemm: row_add needs row_period > 0");
  CGPT_REQUIRE(p.act != CGPT_ACT_SWIGLU || (!p.out_f32 && p.resid == nullptr && p.row_add == nullptr),
               "gemm: SwiGLU epilogue writes bf16 and takes no residual");

  // force_bn: low 12 bits = N tile (0 = auto); 0x1000 forces 1-CTA tiles, 0x2000 forces CTA pairs
  int ctas = M > BM ? 2 : 1;
  if (force_bn & 0x1000) ctas = 1;
  if (force_bn & 0x2000) ctas = 2;
  // measured on B200: CTA pairs are fastest with 256-wide tiles even when N (1408, 4224) leaves a
  // masked tail; 1-CTA tiles prefer an exact divisor of N
  int bn = (force_bn & 0xfff);
  if (bn == 0) bn = ctas == 2 ? (N > 176 ? 256 : (N > 128 ? 176 : 128)) : pick_bn(N);
  if ((force_bn & 0xfff) == 0 && rp == nullptr && e->max_ctas == 0) bn = pick_bn_skinny(M, N, ctas, bn, g_num_sms / ctas);
  if (rp != nullptr) bn = 256;   // the fused rotary epilogue needs one 128-wide head per epilogue warp
  if (p.hm_T > 0) bn = 256;      // the head-major scatter lives in the 256-wide instantiations only
  {
    // Band height.  A weight that fits L2 several times over (ViT / Q-Former linears, <= 24 MB) stays resident
    // whatever the order, so a LOW band (4 m-tiles) is best: the n-tiles that share an A tile then run close
    // together and A is read from HBM once (ViT fc2: 10.6 -> 7.2 GB of DRAM reads, 4.00 -> 3.63 ms per launch, ncu).
    // The big Llama weights (90-180 MB) must instead be amortised over a high band of 16 m-tiles.
    // Sweep: profiles/r01_gemm_band_traffic.txt; CGPT_GEMM_GROUP_M overrides.
    static const int forced = getenv("CGPT_GEMM_GROUP_M") ? atoi(getenv("CGPT_GEMM_GROUP_M")) : 0;
    const long long w_bytes = static_cast<long long>(N) * K * 2;
    p.group_m = forced > 0 ? forced : (w_bytes <= (24LL << 20) ? 4 : GROUP_M_MAX);
  }
  if (p.red_inplace == 3 && !(bn == 256 && ctas == 2 && rp == nullptr && p.hm_T == 0)) p.red_inplace = 2;   // RMW instantiation: CTA pairs, 256-wide tiles
  CUtensorMap ta, tb;
  if (int rc = make_tmap(&ta, A, M, K, lda, BM)) return rc;
  if (int rc = make_tmap(&tb, W, N, K, ldw, bn / ctas)) return rc;
  // timing hook: only outside stream capture (events cannot bracket a graph node)
  GemmProfRec* rec = nullptr;
  if (g_prof_on) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(stream, &cs) == cudaSuccess && cs == cudaStreamCaptureStatusNone) {
      g_prof.push_back(GemmProfRec{nullptr, nullptr, M, N, K});
      rec = &g_prof.back();
      CGPT_CHECK_CUDA(cudaEventCreate(&rec->e0));
      CGPT_CHECK_CUDA(cudaEventCreate(&rec->e1));
      CGPT_CHECK_CUDA(cudaEventRecord(rec->e0, stream));
    }
  }
  int rc = 0;
  // 512 x 256 tiles per CTA pair (two M sub-tiles per CTA sharing the B tile) for the longest K only (Llama down, K = 11008:
  // 1185 -> 1282 TFLOP/s sustained; its DRAM over-read was the worst).  At K = 4096 / 6144 the epilogue this tiling
  // exposes costs more than the saved operand traffic gives back (-4..-7 %, profiles/r02_gemm_vs_cublas.txt); cuBLAS
  // affords the same tiling everywhere with a TMA-store epilogue.  CGPT_GEMM_MT = 1 / 2 forces one form.
  const char* mt_env = getenv("CGPT_GEMM_MT");      // read per call: the tests switch it inside one process
  const int mt_forced = mt_env ? atoi(mt_env) : 0;
  const bool big = ctas == 2 && bn == 256 && M >= 4096 && rp == nullptr && p.hm_T == 0 && p.red_inplace != 3 &&
                   (mt_forced == 2 || (mt_forced == 0 && K >= 8192));   // common epilogues only
  if (big) p.group_m = p.group_m > 1 ? p.group_m / 2 : 1;
  // the 512-row tiles expose their epilogue: take the faster, sector-coalesced form of the in-place residual reductions
  // (llama_down sustained: 1217 -> 1267 TFLOP/s; with 256-row tiles the two forms tie, scripts/gemm_down_probe.py)
  if (big && p.red_inplace == 1 && p.ldo % 4 == 0 && !getenv("CGPT_GEMM_RED_DIRECT")) p.red_inplace = 2;
  if (rp != nullptr) {
    CGPT_REQUIRE(bn == 256, "gemm(rope): needs 256-wide N tiles (got %d)", bn);
    rc = ctas == 2 ? launch_gemm<256, 2, MODE_ROPE>(ta, tb, M, N, K, p, e->max_ctas, stream)
                   : launch_gemm<256, 1, MODE_ROPE>(ta, tb, M, N, K, p, e->max_ctas, stream);
  } else if (p.hm_T > 0) {
    rc = ctas == 2 ? launch_gemm<256, 2, MODE_HM>(ta, tb, M, N, K, p, e->max_ctas, stream)
                   : launch_gemm<256, 1, MODE_HM>(ta, tb, M, N, K, p, e->max_ctas, stream);
  } else if (p.red_inplace == 3) {
    rc = launch_gemm<256, 2, MODE_RMW>(ta, tb, M, N, K, p, e->max_ctas, stream);
  } else if (big) {
    rc = launch_gemm<256, 2, MODE_PLAIN, 2>(ta, tb, M, N, K, p, e->max_ctas, stream);
  } else
  switch (bn * 10 + ctas) {
    case 2561: rc = launch_gemm<256, 1>(ta, tb, M, N, K, p, e->max_ctas, stream); break;
    case 1761: rc = launch_gemm<176, 1>(ta, tb, M, N, K, p, e->max_ctas, stream); break;
    case 1281: rc = launch_gemm<128, 1>(ta, tb, M, N, K, p, e->max_ctas, stream); break;
    case 2562: rc = launch_gemm<256, 2>(ta, tb, M, N, K, p, e->max_ctas, stream); break;
    case 1762: rc = launch_gemm<176, 2>(ta, tb, M, N, K, p, e->max_ctas, stream); break;
    case 1282: rc = launch_gemm<128, 2>(ta, tb, M, N, K, p, e->max_ctas, stream); break;
    default: CGPT_REQUIRE(false, "gemm: unsupported N tile %d", bn);
  }
  if (rec != nullptr && rc == 0) CGPT_CHECK_CUDA(cudaEventRecord(rec->e1, stream));
  return rc;
}

int gemm_profile_begin() {
  for (auto& r : g_prof) {
    if (r.e0) cudaEventDestroy(r.e0);
    if (r.e1) cudaEventDestroy(r.e1);
  }
  g_prof.clear();
  g_prof_on = true;
  return 0;
}

int gemm_profile_end(float* ms_out, int* mnk_out, int capacity, int* count) {
  g_prof_on = false;
  CGPT_CHECK_CUDA(cudaDeviceSynchronize());
  int n = 0;
  for (auto& r : g_prof) {
    if (n < capacity && ms_out != nullptr && mnk_out != nullptr) {
      float ms = 0.f;
      CGPT_CHECK_CUDA(cudaEventElapsedTime(&ms, r.e0, r.e1));
      ms_out[n] = ms;
      mnk_out[3 * n] = r.M; mnk_out[3 * n + 1] = r.N; mnk_out[3 * n + 2] = r.K;
    }
    ++n;
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  g_prof.clear();
  if (count != nullptr) *count = n;
  return 0;
}

}  // namespace cgpt
