// Multi-tile non-causal attention on the 5th-gen tensor cores for sequences that do not fit one UMMA N:
//   EVA ViT self-attention at 448 px, T = 1025 tokens, 16 heads x 88 - the resolution of every shipped reference config
//   (configs/eval_configs/vqav2_eval_noise_0.yaml:35; eva_vit.py:123-153).
// Head-major q, k, v ([B][H][T][hd], see attn_vit.cu).  Work unit = (sample, head, PAIR of 128-query tiles); keys in tiles
// of 128.  Softmax group g (4 warps, thread = query row) owns query tile 2 * pair + g with its own TMEM region
// [S_g 128 columns | O_g hd columns]; BOTH groups consume every K / V tile from shared memory, so a tile is fetched from
// L2 once per 256 query rows.  (The first version gave every 128-row unit its own pass over K and V: 32 GB of L2 -> SM
// reads per layer at batch 256 = 6.1 TB/s, the chip's L2 bandwidth, and 5.3 ms - profiles/r02_ncu_kernels_summary.txt.)
// All of S (1025 columns) does not fit TMEM, so softmax is EXACT two-pass: pass 1 forms S = Q K^T tile by tile only for
// the row maximum, pass 2 forms it again, turns it into P = exp2(s - max) in place (packed bf16, read back by the tensor
// core as the A operand of P.V) and accumulates O += P V in TMEM across the key tiles: no online rescale of O.
//   warp 0      TMA producer : the pair's two Q tiles, then the K / V tiles in consumption order through a 4-slot ring
//   warp 1      MMA issuer   : per key tile S_0, S_1 (pass 1, 2) and P_0.V, P_1.V (pass 2)
//   warps 4-7   softmax group 0, warps 8-11 softmax group 1
// The cls token is simply row 0 of the head block: tiles start at the block's first row, keys >= T are masked, rows >= T
// are not stored (T = 1025 = 8 full tiles + 1 row; the specialised kernel in attn_vit.cu folds the cls row in on CUDA
// cores instead, which pays off at T = 257 where a third tile would cost 2.25x).
#include <stdlib.h>
#include "common.cuh"
#include "ops.h"

namespace cgpt {

struct LongAttnParams {
  __nv_bfloat16* o; long long ldo;   // row-major [B*T, >= H*hd]
  int H, hd, hd16, T;
  float scale_log2e;
  int n_tiles;                       // ceil(T / 128): key tiles = query tiles
  int n_pairs;                       // ceil(n_tiles / 2): query-tile pairs per (sample, head)
  int n_units;                       // B * H * n_pairs
  int pv_n;
  int last_cols;                     // valid keys of the last key tile, rounded up to 16: its UMMA N and P.V depth
};

constexpr int LA_THREADS = 384;
constexpr int LA_SUB = 16384;        // 128 rows x 128 B (one swizzle atom wide)
constexpr int LA_TILE = 2 * LA_SUB;  // [cols 0..63 | cols 64..]
constexpr int LA_STAGES = 4;
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
  uint8_t* sQ = smem;                               // two Q tiles (one per softmax group)
  uint8_t* ring = smem + 2 * LA_TILE;               // LA_STAGES K / V tiles
  uint8_t* tail = ring + LA_STAGES * LA_TILE;
  // every barrier is used in the same order by all its parties: use k has parity k & 1
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail);
  uint64_t* full = bars;                            // [STAGES] TMA -> MMA
  uint64_t* empty = bars + LA_STAGES;               // [STAGES] MMA -> TMA
  uint64_t* bar_q = bars + 2 * LA_STAGES;           // both Q tiles landed
  uint64_t* bar_qfree = bar_q + 1;                  // the unit's last S products retired: Q slots free
  uint64_t* bar_s = bar_q + 2;                      // [2] S_g formed                     (once per tile of pass 1 and pass 2)
  uint64_t* bar_sr = bar_q + 4;                     // [2] pass 1: S_g read by its group  (128 arrivals, once per pass-1 tile)
  uint64_t* bar_p = bar_q + 6;                      // [2] pass 2: P_g written            (128 arrivals, once per pass-2 tile)
  uint64_t* bar_o = bar_q + 8;                      // O of both groups complete (once per unit)
  uint64_t* bar_ofree = bar_q + 9;                  // O drained by both groups (256 arrivals, once per unit)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar_q + 10);
  uint8_t* ostage = tail + 256;                     // [8 softmax warps][32 rows][LA_OPITCH]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int hd = p.hd, T = p.T, n = p.n_tiles, np = p.n_pairs;
  const int ksteps_s = p.hd16 / 16;
  const int grid = static_cast<int>(gridDim.x);
  const int first_unit = static_cast<int>(blockIdx.x);
  const uint32_t tile_tx = 128u * static_cast<uint32_t>(p.hd16) * 2u;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_q0); tma_prefetch_desc(&map_q1); tma_prefetch_desc(&map_k0);
    tma_prefetch_desc(&map_k1); tma_prefetch_desc(&map_v0); tma_prefetch_desc(&map_v1);
    for (int i = 0; i < LA_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(bar_q, 1); mbar_init(bar_qfree, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_s[i], 1); mbar_init(&bar_sr[i], 128); mbar_init(&bar_p[i], 128);
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
  // TMEM region of group g: S_g (later P_g in its first 64 columns) at g * 256, O_g at g * 256 + 128

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
        const int item = u / np, pr = u - item * np;
        const int row0 = item * T;
        const int early = n < 2 ? n : 2;   // the first K tiles do not have to wait for the Q slots
        for (int kt = 0; kt < early; ++kt) push(&map_k0, &map_k1, row0 + kt * 128);
        if (ui > 0) mbar_wait(bar_qfree, (ui - 1) & 1);
        mbar_arrive_expect_tx(bar_q, 2 * tile_tx);
        for (int g = 0; g < 2; ++g) {     // a second tile past the head block reads finite neighbour rows, never stored
          tma_load_2d(sQ + g * LA_TILE, &map_q0, bar_q, 0, row0 + (2 * pr + g) * 128);
          tma_load_2d(sQ + g * LA_TILE + LA_SUB, &map_q1, bar_q, 64, row0 + (2 * pr + g) * 128);
        }
        for (int kt = early; kt < n; ++kt) push(&map_k0, &map_k1, row0 + kt * 128);      // pass 1: K
        for (int kt = 0; kt < n; ++kt) {                                                  // pass 2: K_kt, V_kt
          push(&map_k0, &map_k1, row0 + kt * 128);
          push(&map_v0, &map_v1, row0 + kt * 128);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (one thread)
    // tcgen05.mma instructions of one thread execute in issue order, so an S product issued after the P.V that read the
    // same TMEM columns needs no barrier; only the softmax groups' tcgen05.ld / st are waited for (bar_sr, bar_p).
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc_bf16(128, 128);
      const uint32_t idesc_s_last = make_idesc_bf16(128, p.last_cols);
      const uint32_t idesc_o = make_idesc_bf16(128, p.pv_n) | (1u << 16);   // B (= V) is MN-major
      int stage = 0;
      uint32_t phase = 0;
      int n_sr[2] = {0, 0}, n_p[2] = {0, 0};   // per group: pass-1 tiles issued, pass-2 tiles whose P has been consumed
      auto advance = [&]() { if (++stage == LA_STAGES) { stage = 0; phase ^= 1; } };
      auto issue_s = [&](int g, const uint8_t* sK, bool last_tile) {
        for (int ks = 0; ks < ksteps_s; ++ks) {
          const uint64_t ad = make_smem_desc_sw128(smem_u32(sQ + g * LA_TILE + (ks >> 2) * LA_SUB)) + 2 * (ks & 3);
          const uint64_t bd = make_smem_desc_sw128(smem_u32(sK + (ks >> 2) * LA_SUB)) + 2 * (ks & 3);
          umma_bf16(tmem_base + g * 256, ad, bd, last_tile ? idesc_s_last : idesc_s, ks != 0);
        }
        umma_commit(&bar_s[g]);
      };
      int ui = 0;
      for (int u = first_unit; u < p.n_units; u += grid, ++ui) {
        const int pr = u % np;
        const int ng = (2 * pr + 1 < n) ? 2 : 1;      // the last pair of an odd tile count has one query tile
        mbar_wait(bar_q, ui & 1);
        // ---- pass 1: S only (row maximum)
        for (int kt = 0; kt < n; ++kt) {
          mbar_wait(&full[stage], phase);
          const uint8_t* sK = ring + stage * LA_TILE;
          for (int g = 0; g < ng; ++g) {
            if (kt > 0) mbar_wait(&bar_sr[g], (n_sr[g] - 1) & 1);   // the group has read the previous tile's scores
            tcgen05_fence_after();
            issue_s(g, sK, kt == n - 1);
            ++n_sr[g];
          }
          umma_commit(&empty[stage]);
          advance();
        }
        // ---- pass 2: S_g(0); then per key tile P_g(kt).V followed at once by S_g(kt + 1)
        mbar_wait(&full[stage], phase);
        for (int g = 0; g < ng; ++g) {
          mbar_wait(&bar_sr[g], (n_sr[g] - 1) & 1);
          tcgen05_fence_after();
          issue_s(g, ring + stage * LA_TILE, n == 1);
        }
        if (n == 1) umma_commit(bar_qfree);
        umma_commit(&empty[stage]);
        advance();
        for (int kt = 0; kt < n; ++kt) {
          const bool more = kt + 1 < n;
          const int sv = stage;
          mbar_wait(&full[sv], phase);                // V_kt
          advance();
          const int sk = stage;
          if (more) { mbar_wait(&full[sk], phase); advance(); }   // K_{kt+1}
          if (kt == 0 && ui > 0) mbar_wait(bar_ofree, (ui - 1) & 1);   // the previous unit's O has been drained
          const uint8_t* sV = ring + sv * LA_TILE;
          const int pv_steps = (kt == n - 1) ? p.last_cols / 16 : 8;
          for (int g = 0; g < ng; ++g) {
            mbar_wait(&bar_p[g], n_p[g] & 1);
            ++n_p[g];
            tcgen05_fence_after();
            for (int ks = 0; ks < pv_steps; ++ks) {
              const uint64_t bd = la_desc_mn(smem_u32(sV + ks * 16 * 128), LA_SUB);
              umma_bf16_ts(tmem_base + g * 256 + 128, tmem_base + g * 256 + ks * 8, bd, idesc_o, (kt | ks) != 0);
            }
            if (g == ng - 1) umma_commit(&empty[sv]);
            if (more) issue_s(g, ring + sk * LA_TILE, kt + 1 == n - 1);
          }
          if (more) {
            if (kt + 1 == n - 1) umma_commit(bar_qfree);   // the unit's last read of the Q tiles
            umma_commit(&empty[sk]);
          }
        }
        umma_commit(bar_o);
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ softmax groups, thread = query row
    const int g = (warp - 4) >> 2;                    // group g owns query tile 2 * pair + g
    const uint32_t t_s = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + g * 256;
    const uint32_t t_o = t_s + 128;
    uint8_t* patch = ostage + (warp - 4) * (32 * LA_OPITCH);
    int cnt_s = 0;                                    // uses of bar_s[g] so far
    int ui = 0;
    for (int u = first_unit; u < p.n_units; u += grid, ++ui) {
      const int item = u / np, pr = u - item * np;
      const int qt = 2 * pr + g;
      const int h = item % p.H, bb = item / p.H;
      if (qt >= n) {                                  // odd tile count: the last pair has no second query tile
        mbar_arrive(bar_ofree);
        continue;
      }
      // ---- pass 1: row maximum
      float mx = -INFINITY;
      for (int kt = 0; kt < n; ++kt) {
        mbar_wait(&bar_s[g], cnt_s & 1);
        ++cnt_s;
        tcgen05_fence_after();
        const int kmax = T - kt * 128;                // keys [0, kmax) of this tile exist
        const int ncols = kt == n - 1 ? p.last_cols : 128;   // columns the tensor core wrote
        uint32_t va[16], vb[16];
        tmem_ld_x16(t_s, va);
#pragma unroll 1
        for (int c = 0; c < ncols; c += 32) {
          tmem_ld_wait();
          if (c + 16 < ncols) tmem_ld_x16(t_s + c + 16, vb);
          if (c + 16 <= kmax) {
#pragma unroll
            for (int i = 0; i < 16; ++i) mx = fmaxf(mx, __uint_as_float(va[i]));
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) if (c + i < kmax) mx = fmaxf(mx, __uint_as_float(va[i]));
          }
          if (c + 16 >= ncols) break;
          tmem_ld_wait();
          if (c + 32 < ncols) tmem_ld_x16(t_s + c + 32, va);
          if (c + 32 <= kmax) {
#pragma unroll
            for (int i = 0; i < 16; ++i) mx = fmaxf(mx, __uint_as_float(vb[i]));
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) if (c + 16 + i < kmax) mx = fmaxf(mx, __uint_as_float(vb[i]));
          }
        }
        tcgen05_fence_before();
        mbar_arrive(&bar_sr[g]);
      }
      if (mx == -INFINITY) mx = 0.f;
      const float neg_ms = -mx * p.scale_log2e;
      // ---- pass 2: P = exp2(s * scale - max * scale), packed bf16 in place over the first 64 columns of S_g
      float sum = 0.f;
      for (int kt = 0; kt < n; ++kt) {
        mbar_wait(&bar_s[g], cnt_s & 1);
        ++cnt_s;
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
        const int ncols = kt == n - 1 ? p.last_cols : 128;
        uint32_t va[16], vb[16];
        tmem_ld_x16(t_s, va);
#pragma unroll 1
        for (int c = 0; c < ncols; c += 32) {
          tmem_ld_wait();
          if (c + 16 < ncols) tmem_ld_x16(t_s + c + 16, vb);
          emit(va, c);
          if (c + 16 >= ncols) break;
          tmem_ld_wait();
          if (c + 32 < ncols) tmem_ld_x16(t_s + c + 32, va);
          emit(vb, c + 16);
        }
        tmem_st_wait();
        tcgen05_fence_before();
        mbar_arrive(&bar_p[g]);
      }
      // ---- epilogue: O_g / rowsum -> bf16 -> shared-memory transpose -> global
      mbar_wait(bar_o, ui & 1);
      tcgen05_fence_after();
      const float inv = sum > 0.f ? 1.f / sum : 0.f;
      const int q_warp0 = qt * 128 + (warp & 3) * 32;
      __nv_bfloat16* obase = p.o + (static_cast<long long>(bb) * T + q_warp0) * p.ldo + h * hd;
#pragma unroll 1
      for (int c = 0; c < hd; c += 32) {
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
  if (a->B * (long long)a->H * ((a->Tk + 255) / 256) > 0x7fffffffLL) return 0;
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
  p.n_pairs = (p.n_tiles + 1) / 2;
  p.n_units = a->B * a->H * p.n_pairs;
  p.pv_n = p.hd16;
  p.last_cols = ((a->Tk - (p.n_tiles - 1) * 128) + 15) & ~15;
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const int smem = (2 + LA_STAGES) * LA_TILE + 256 + 8 * 32 * LA_OPITCH + 1024;
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
