// Multi-tile non-causal attention on the 5th-gen tensor cores for sequences that do not fit one UMMA N:
//   EVA ViT self-attention at 448 px, T = 1025 tokens, 16 heads x 88 - the resolution of every shipped reference config
//   (configs/eval_configs/vqav2_eval_noise_0.yaml:35; eva_vit.py:123-153).
// Head-major q, k, v ([B][H][T][hd], see attn_vit.cu).  Work unit = (sample, head, 128-query tile); keys in tiles of 128.
// TMEM holds THREE S buffers of 128 x 128 fp32 (tile number modulo 3) and the O accumulator (128 x hd) - all of S (1025
// columns) would not fit - so softmax is EXACT two-pass: pass 1 forms S = Q K^T tile by tile only for the row maximum, pass 2 forms it again,
// turns it into P = exp2(s - max) in place (packed bf16, read back by the tensor core as the A operand of P.V) and
// accumulates O += P V in TMEM across the key tiles: no online rescale of O, and the tensor pipe has the slack for the
// second Q K^T (the kernel is MUFU-bound: 1025 exp2 per row).
//   warp 0      TMA producer : Q tile, then the K / V tiles in consumption order through a 5-slot ring
//   warp 1      MMA issuer   : S two tiles ahead of P.V (three S buffers), so a softmax group never waits for the tensor
//                              core: with two buffers every S -> softmax -> P.V -> S hand-over (4 mbarrier round trips
//                              per tile) sat on the critical path: 39k cycles per unit against 14k of MUFU + tensor work
//   warps 4-7   softmax group 0: the EVEN key tiles;  warps 8-11  softmax group 1: the ODD key tiles (thread = query row);
//               the two groups exchange row maximum and row sum through shared memory
// The cls token is simply row 0 of the head block: tiles start at the block's first row, keys >= T are masked, rows >= T
// are not stored (T = 1025 = 8 full tiles + 1 row; the specialised kernel in attn_vit.cu folds the cls row in on CUDA
// cores instead, which pays off at T = 257 where a 129th tile would cost 2.25x).
#include <stdlib.h>
#include "common.cuh"
#include "ops.h"

namespace cgpt {

struct LongAttnParams {
  __nv_bfloat16* o; long long ldo;   // row-major [B*T, >= H*hd]
  int H, hd, hd16, T;
  float scale_log2e;
  int n_tiles;                       // ceil(T / 128): key tiles = query tiles
  int n_units;                       // B * H * n_tiles
  int pv_n;
};

constexpr int LA_THREADS = 384;
constexpr int LA_SUB = 16384;        // 128 rows x 128 B (one swizzle atom wide)
constexpr int LA_TILE = 2 * LA_SUB;  // [cols 0..63 | cols 64..]
constexpr int LA_STAGES = 5;
constexpr int LA_OPITCH = 80;

__device__ __forceinline__ uint64_t la_desc_mn(uint32_t addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>(1024u >> 4) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}
__device__ __forceinline__ float la_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(LA_THREADS, 1)
attn_long_kernel(const __grid_constant__ CUtensorMap map_q0, const __grid_constant__ CUtensorMap map_q1,
                 const __grid_constant__ CUtensorMap map_k0, const __grid_constant__ CUtensorMap map_k1,
                 const __grid_constant__ CUtensorMap map_v0, const __grid_constant__ CUtensorMap map_v1,
                 LongAttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;                               // one Q tile
  uint8_t* ring = smem + LA_TILE;                   // LA_STAGES K / V tiles
  uint8_t* tail = ring + LA_STAGES * LA_TILE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail);
  uint64_t* full = bars;                            // [STAGES] TMA -> MMA
  uint64_t* empty = bars + LA_STAGES;               // [STAGES] MMA -> TMA
  uint64_t* bar_q = bars + 2 * LA_STAGES;           // Q tile landed
  uint64_t* bar_qfree = bar_q + 1;                  // the unit's last S product retired: Q slot free
  uint64_t* bar_s = bar_q + 2;                      // [3] S in buffer b
  uint64_t* bar_sr = bar_q + 5;                     // [3] pass 1: buffer b read by its softmax group (128 arrivals)
  uint64_t* bar_p = bar_q + 8;                      // [3] pass 2: P in buffer b (128 arrivals)
  uint64_t* bar_pv = bar_q + 11;                    // [3] P.V out of buffer b retired
  uint64_t* bar_o = bar_q + 14;                     // O of the unit complete
  uint64_t* bar_ofree = bar_q + 15;                 // O drained by both groups (256 arrivals)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar_q + 16);
  float* xch = reinterpret_cast<float*>(tail + 256);                 // [2 unit parities][2 groups][max | sum][128]
  uint8_t* ostage = reinterpret_cast<uint8_t*>(xch + 2 * 2 * 2 * 128);   // [8 softmax warps][32 rows][LA_OPITCH]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int hd = p.hd, T = p.T, n = p.n_tiles;
  const int ksteps_s = p.hd16 / 16;
  const int grid = static_cast<int>(gridDim.x);
  const int first_unit = static_cast<int>(blockIdx.x);
  const uint32_t tile_tx = 128u * static_cast<uint32_t>(p.hd16) * 2u;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_q0); tma_prefetch_desc(&map_q1); tma_prefetch_desc(&map_k0);
    tma_prefetch_desc(&map_k1); tma_prefetch_desc(&map_v0); tma_prefetch_desc(&map_v1);
    for (int i = 0; i < LA_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(bar_q, 1); mbar_init(bar_qfree, 1);
    for (int i = 0; i < 3; ++i) {
      mbar_init(&bar_s[i], 1); mbar_init(&bar_sr[i], 128); mbar_init(&bar_p[i], 128); mbar_init(&bar_pv[i], 1);
    }
    mbar_init(bar_o, 1); mbar_init(bar_ofree, 256);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  constexpr uint32_t O_COL = 384;   // S buffers at columns 0, 128, 256

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      auto push = [&](const CUtensorMap* m0, const CUtensorMap* m1, int row) {
        mbar_wait(&empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&full[stage], tile_tx);
        uint8_t* dst = ring + stage * LA_TILE;
        tma_load_2d(dst, m0, &full[stage], 0, row);
        tma_load_2d(dst + LA_SUB, m1, &full[stage], 64, row);
        if (++stage == LA_STAGES) { stage = 0; phase ^= 1; }
      };
      int ui = 0;
      for (int u = first_unit; u < p.n_units; u += grid, ++ui) {
        const int item = u / n, qt = u - item * n;
        const int row0 = item * T;
        if (ui > 0) mbar_wait(bar_qfree, (ui - 1) & 1);
        mbar_arrive_expect_tx(bar_q, tile_tx);
        tma_load_2d(sQ, &map_q0, bar_q, 0, row0 + qt * 128);
        tma_load_2d(sQ + LA_SUB, &map_q1, bar_q, 64, row0 + qt * 128);
        for (int kt = 0; kt < n; ++kt) push(&map_k0, &map_k1, row0 + kt * 128);          // pass 1
        for (int kt = 0; kt < n; ++kt) {                                                  // pass 2: K_kt, then V_{kt-2}
          push(&map_k0, &map_k1, row0 + kt * 128);
          if (kt >= 2) push(&map_v0, &map_v1, row0 + (kt - 2) * 128);
        }
        if (n >= 2) push(&map_v0, &map_v1, row0 + (n - 2) * 128);
        push(&map_v0, &map_v1, row0 + (n - 1) * 128);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (one thread)
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc_bf16(128, 128);
      const uint32_t idesc_o = make_idesc_bf16(128, p.pv_n) | (1u << 16);   // B (= V) is MN-major
      int stage = 0;
      uint32_t phase = 0;
      int n_s1[3] = {0, 0, 0}, n_pv[3] = {0, 0, 0};   // pass-1 uses / P.V uses of each S buffer so far
      int last_kind[3] = {0, 0, 0};                  // 0 = never used, 1 = pass-1 tile, 2 = pass-2 tile
      auto wait_free = [&](int b) {
        if (last_kind[b] == 1) mbar_wait(&bar_sr[b], (n_s1[b] - 1) & 1);
        else if (last_kind[b] == 2) mbar_wait(&bar_pv[b], (n_pv[b] - 1) & 1);
      };
      auto issue_s = [&](int b) {
        mbar_wait(&full[stage], phase);
        tcgen05_fence_after();
        const uint8_t* sK = ring + stage * LA_TILE;
        for (int ks = 0; ks < ksteps_s; ++ks) {
          const uint64_t ad = make_smem_desc_sw128(smem_u32(sQ + (ks >> 2) * LA_SUB)) + 2 * (ks & 3);
          const uint64_t bd = make_smem_desc_sw128(smem_u32(sK + (ks >> 2) * LA_SUB)) + 2 * (ks & 3);
          umma_bf16(tmem_base + b * 128, ad, bd, idesc_s, ks != 0);
        }
        umma_commit(&empty[stage]);
        umma_commit(&bar_s[b]);
        if (++stage == LA_STAGES) { stage = 0; phase ^= 1; }
      };
      int ui = 0;
      long long seq = 0;                             // S tiles issued so far (both passes, all units): buffer = seq % 3
      for (int u = first_unit; u < p.n_units; u += grid, ++ui) {
        const long long seq2 = seq + n;              // sequence number of this unit's first pass-2 tile
        auto issue_pv = [&](int j) {
          const int b = static_cast<int>((seq2 + j) % 3);
          mbar_wait(&bar_p[b], (n_pv[b]) & 1);
          mbar_wait(&full[stage], phase);
          if (j == 0 && ui > 0) mbar_wait(bar_ofree, (ui - 1) & 1);     // the previous unit's O has been drained
          tcgen05_fence_after();
          const uint8_t* sV = ring + stage * LA_TILE;
          for (int ks = 0; ks < 8; ++ks) {
            const uint64_t bd = la_desc_mn(smem_u32(sV + ks * 16 * 128), LA_SUB);
            umma_bf16_ts(tmem_base + O_COL, tmem_base + b * 128 + ks * 8, bd, idesc_o, (j | ks) != 0);
          }
          umma_commit(&empty[stage]);
          umma_commit(&bar_pv[b]);
          ++n_pv[b];
          if (++stage == LA_STAGES) { stage = 0; phase ^= 1; }
        };
        mbar_wait(bar_q, ui & 1);
        for (int kt = 0; kt < n; ++kt, ++seq) {       // pass 1: S only (row maximum)
          const int b = static_cast<int>(seq % 3);
          wait_free(b);
          issue_s(b);
          ++n_s1[b];
          last_kind[b] = 1;
        }
        for (int kt = 0; kt < n; ++kt, ++seq) {       // pass 2: S again, P.V two tiles behind
          const int b = static_cast<int>(seq % 3);
          wait_free(b);
          issue_s(b);
          last_kind[b] = 2;
          if (kt == n - 1) umma_commit(bar_qfree);    // the unit's last read of the Q tile
          if (kt >= 2) issue_pv(kt - 2);
        }
        if (n >= 2) issue_pv(n - 2);
        issue_pv(n - 1);
        umma_commit(bar_o);
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ softmax groups, thread = query row
    const int g = (warp - 4) >> 2;                    // group g owns the key tiles kt = g, g + 2, ...
    const int r = (warp & 3) * 32 + lane;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    const uint32_t t_o = t_lane + O_COL;
    uint8_t* patch = ostage + (warp - 4) * (32 * LA_OPITCH);
    int ui = 0;
    for (int u = first_unit; u < p.n_units; u += grid, ++ui) {
      const int item = u / n, qt = u - item * n;
      const int h = item % p.H, bb = item / p.H;
      float* xmax = xch + (ui & 1) * 512;             // [2 groups][128]
      float* xsum = xmax + 256;
      // ---- pass 1: row maximum over this group's key tiles
      // S tile number of this unit's pass-1 tile 0 (2n tiles per unit): buffer = number % 3, its use number / 3
      const long long seq1 = static_cast<long long>(ui) * 2 * n;
      float mx = -INFINITY;
      for (int kt = g; kt < n; kt += 2) {
        const long long sq = seq1 + kt;
        const int sb = static_cast<int>(sq % 3);
        const uint32_t t_s = t_lane + sb * 128;
        mbar_wait(&bar_s[sb], static_cast<uint32_t>((sq / 3) & 1));
        tcgen05_fence_after();
        const int kmax = T - kt * 128;                // keys [0, kmax) of this tile exist
        uint32_t va[16], vb[16];
        tmem_ld_x16(t_s, va);
#pragma unroll 1
        for (int c = 0; c < 128; c += 32) {
          tmem_ld_wait();
          tmem_ld_x16(t_s + c + 16, vb);
          if (c + 16 <= kmax) {
#pragma unroll
            for (int i = 0; i < 16; ++i) mx = fmaxf(mx, __uint_as_float(va[i]));
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) if (c + i < kmax) mx = fmaxf(mx, __uint_as_float(va[i]));
          }
          tmem_ld_wait();
          if (c + 32 < 128) tmem_ld_x16(t_s + c + 32, va);
          if (c + 32 <= kmax) {
#pragma unroll
            for (int i = 0; i < 16; ++i) mx = fmaxf(mx, __uint_as_float(vb[i]));
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) if (c + 16 + i < kmax) mx = fmaxf(mx, __uint_as_float(vb[i]));
          }
        }
        tcgen05_fence_before();
        mbar_arrive(&bar_sr[sb]);
      }
      xmax[g * 128 + r] = mx;
      asm volatile("bar.sync 3, 256;" ::: "memory");
      mx = fmaxf(xmax[r], xmax[128 + r]);
      if (mx == -INFINITY) mx = 0.f;
      const float neg_ms = -mx * p.scale_log2e;
      // ---- pass 2: P = exp2(s * scale - max * scale), packed bf16 in place over the first 64 columns of the buffer
      float sum = 0.f;
      for (int kt = g; kt < n; kt += 2) {
        const long long sq = seq1 + n + kt;
        const int sb = static_cast<int>(sq % 3);
        const uint32_t t_s = t_lane + sb * 128;
        mbar_wait(&bar_s[sb], static_cast<uint32_t>((sq / 3) & 1));
        tcgen05_fence_after();
        const int kmax = T - kt * 128;
        auto emit = [&](const uint32_t* v, int c) {
          float e[16];
          if (c + 16 <= kmax) {
#pragma unroll
            for (int i = 0; i < 16; ++i) e[i] = la_exp2(fmaf(__uint_as_float(v[i]), p.scale_log2e, neg_ms));
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              e[i] = (c + i < kmax) ? la_exp2(fmaf(__uint_as_float(v[i]), p.scale_log2e, neg_ms)) : 0.f;
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) sum += e[i];
          uint32_t w[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) w[i] = pack_bf16x2(e[2 * i], e[2 * i + 1]);
          tmem_st_x8(t_s + (c >> 1), w);
        };
        uint32_t va[16], vb[16];
        tmem_ld_x16(t_s, va);
#pragma unroll 1
        for (int c = 0; c < 128; c += 32) {
          tmem_ld_wait();
          tmem_ld_x16(t_s + c + 16, vb);
          emit(va, c);
          tmem_ld_wait();
          if (c + 32 < 128) tmem_ld_x16(t_s + c + 32, va);
          emit(vb, c + 16);
        }
        tmem_st_wait();
        tcgen05_fence_before();
        mbar_arrive(&bar_p[sb]);
      }
      xsum[g * 128 + r] = sum;
      asm volatile("bar.sync 3, 256;" ::: "memory");
      sum = xsum[r] + xsum[128 + r];
      // ---- epilogue: O / rowsum -> bf16 -> shared-memory transpose -> global; 32-column groups alternate between the groups
      mbar_wait(bar_o, ui & 1);
      tcgen05_fence_after();
      const float inv = sum > 0.f ? 1.f / sum : 0.f;
      const int q_warp0 = qt * 128 + (warp & 3) * 32;
      __nv_bfloat16* obase = p.o + (static_cast<long long>(bb) * T + q_warp0) * p.ldo + h * hd;
      int grp = 0;
#pragma unroll 1
      for (int c = 0; c < hd; c += 32, ++grp) {
        if ((grp & 1) != g) continue;                 // warp-uniform
        uint32_t v[32];
        tmem_ld_x32(t_o + c, v);
        tmem_ld_wait();
        uint32_t w[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2)
          w[i >> 1] = pack_bf16x2(__uint_as_float(v[i]) * inv, __uint_as_float(v[i + 1]) * inv);
        uint4* mine = reinterpret_cast<uint4*>(patch + lane * LA_OPITCH);
        mine[0] = make_uint4(w[0], w[1], w[2], w[3]);
        mine[1] = make_uint4(w[4], w[5], w[6], w[7]);
        mine[2] = make_uint4(w[8], w[9], w[10], w[11]);
        mine[3] = make_uint4(w[12], w[13], w[14], w[15]);
        __syncwarp();
        const int seg = lane & 3;
        if (c + seg * 8 < hd) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int rr = k * 8 + (lane >> 2);
            if (q_warp0 + rr < T)
              *reinterpret_cast<uint4*>(obase + static_cast<long long>(rr) * p.ldo + c + seg * 8) =
                  *reinterpret_cast<const uint4*>(patch + rr * LA_OPITCH + seg * 16);
          }
        }
        __syncwarp();
      }
      tcgen05_fence_before();
      mbar_arrive(bar_ofree);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn4)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn4 la_encode_fn() {
  static EncodeTiledFn4 fn = nullptr;
  if (fn) return fn;
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn4>(ptr);
  return fn;
}
static int la_make_map(CUtensorMap* map, const void* base, long long rows, int cols, int box_cols) {
  EncodeTiledFn4 fn = la_encode_fn();
  CGPT_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled driver entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, 128u};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CGPT_REQUIRE(r == CUDA_SUCCESS, "attention_long: cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%d box=%d", (int)r, rows,
               cols, box_cols);
  return 0;
}

// 1 = this head-major problem is served by the multi-tile tcgen05 kernel
int attn_long_supported(const cgpt_attn_args* a) {
  if (!a->head_major || a->causal || a->P != 0 || a->decode_kernel != 0) return 0;
  if (a->head_dim <= 64 || a->head_dim > 128 || (a->head_dim & 7)) return 0;
  if (a->Tq != a->Tk || a->q_rows_per_batch != a->Tq || a->kv_rows_per_batch != a->Tk || a->Tk < 1) return 0;
  if (a->B * (long long)a->H * a->Tk > 0x7fffffffLL) return 0;
  if (a->B * (long long)a->H * ((a->Tk + 127) / 128) > 0x7fffffffLL) return 0;
  if ((reinterpret_cast<uintptr_t>(a->q) | reinterpret_cast<uintptr_t>(a->k) | reinterpret_cast<uintptr_t>(a->v) |
       reinterpret_cast<uintptr_t>(a->o)) & 15)
    return 0;
  return 1;
}

int attention_long(const cgpt_attn_args* a, cudaStream_t stream) {
  CGPT_REQUIRE(attn_long_supported(a), "attention_long: unsupported head-major problem (Tq=%d Tk=%d hd=%d causal=%d)", a->Tq,
               a->Tk, a->head_dim, a->causal);
  LongAttnParams p;
  p.o = (__nv_bfloat16*)a->o; p.ldo = a->ldo;
  p.H = a->H; p.hd = a->head_dim; p.hd16 = (a->head_dim + 15) & ~15; p.T = a->Tk;
  p.scale_log2e = a->scale * 1.4426950408889634f;
  p.n_tiles = (a->Tk + 127) / 128;
  p.n_units = a->B * a->H * p.n_tiles;
  p.pv_n = p.hd16;
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const int smem = (1 + LA_STAGES) * LA_TILE + 256 + 2 * 2 * 2 * 128 * 4 + 8 * 32 * LA_OPITCH + 1024;
  const long long rows = (long long)a->B * a->H * a->Tk;
  const int c1 = p.hd16 - 64;
  CUtensorMap mq0, mq1, mk0, mk1, mv0, mv1;
  if (int rc = la_make_map(&mq0, a->q, rows, a->head_dim, 64)) return rc;
  if (int rc = la_make_map(&mq1, a->q, rows, a->head_dim, c1)) return rc;
  if (int rc = la_make_map(&mk0, a->k, rows, a->head_dim, 64)) return rc;
  if (int rc = la_make_map(&mk1, a->k, rows, a->head_dim, c1)) return rc;
  if (int rc = la_make_map(&mv0, a->v, rows, a->head_dim, 64)) return rc;
  if (int rc = la_make_map(&mv1, a->v, rows, a->head_dim, c1)) return rc;
  static bool configured = false;
  if (!configured) {
    CGPT_CHECK_CUDA(cudaFuncSetAttribute(attn_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  int grid = sms;
  if (grid > p.n_units) grid = p.n_units;
  attn_long_kernel<<<grid, LA_THREADS, smem, stream>>>(mq0, mq1, mk0, mk1, mv0, mv1, p);
  CGPT_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace cgpt
