// Multi-tile non-causal attention on the 5th-gen tensor cores for sequences that do not fit one UMMA N:
//   EVA ViT self-attention at 448 px, T = 1025 tokens, 16 heads x 88 - the resolution of every shipped reference config
//   (configs/eval_configs/vqav2_eval_noise_0.yaml:35; eva_vit.py:123-153).
// Head-major q, k, v ([B][H][T][hd], see attn_vit.cu).  Work unit = (sample, head, PAIR of 128-query tiles); keys in tiles
// of 128.  Softmax group g (4 warps, thread = query row) owns query tile 2 * pair + g with its own TMEM region
// [S_g 128 columns | O_g 128 columns]; BOTH groups consume every K / V tile from shared memory, so a tile is fetched from
// L2 once per 256 query rows (the first version gave every 128-row unit its own pass over K and V: 32 GB of L2 -> SM
// reads per layer at batch 256; now 13.7 GB - profiles/r02_ncu_attention_final.txt).
// All of S (1025 columns) does not fit TMEM.  The first version ran an exact two-pass softmax (scores formed and read
// twice); this one forms and reads every score once: ONLINE softmax in 64-key steps with the lazy rescale of
// FlashAttention-4: a row keeps a reference maximum m; a step whose maximum exceeds m by more than 8 (log2 units) in some
// row of the warp multiplies that warp's rows of O (TMEM) and the running sums by 2^(m - m_new) first; otherwise
// P = exp2(s - m) <= 2^8 is used as is - bf16 keeps its relative precision, the final O / sum is the same quantity.
// P replaces the scores in place (packed bf16) and is read by the tensor core as the A operand of P.V; O accumulates in
// TMEM across the key tiles.  History of the kernel, 5.29 -> 2.30 ms per layer: profiles/r02_attn_long_probe.txt.
//   warp 0      TMA producer : the pair's two Q tiles, then K_0, V_0, K_1, V_1, ... through a 4-slot ring
//   warps 1, 2  MMA issuers  : one thread per softmax group; per key tile P_h0.V, S_h0(next), P_h1.V, S_h1(next): while
//               the group exponentiates one 64-key half, the tensor pipe consumes the other and refills it
//   warps 3-6   softmax group 0, warps 7-10 softmax group 1 (a warp reaches TMEM lanes 32 * (warp % 4) ..: rows follow)
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
  long long* dbg;                    // optional [grid][16] cycle stamps of each CTA's second unit (CGPT_ATTN_DBG)
};
// debug timers (CGPT_ATTN_DBG): a role sets `dbg_on` (warp-uniform, lane 0 writes) and keeps its sums in registers (`dacc`), written
// out once at the end of the CTA's second unit - a global read-modify-write per sample would stall the thread it measures
#define LA_STAMP(slot) do { if (DBG && dbg_on && ui == 1) { const long long t_ = clock64(); if (lane == 0) p.dbg[blockIdx.x * 16 + (slot)] = t_; } } while (0)
#define LA_TIMED_WAIT(k, bar, par) do { \
    if (DBG && dbg_on && ui == 1) { const long long t0_ = clock64(); mbar_wait(bar, par); dacc[k] += clock64() - t0_; } \
    else mbar_wait(bar, par); } while (0)
#define LA_T0() long long tseg_ = (DBG && dbg_on && ui == 1) ? clock64() : 0
#define LA_SEG(k) do { if (DBG && dbg_on && ui == 1) { const long long t1_ = clock64(); dacc[k] += t1_ - tseg_; tseg_ = t1_; } } while (0)
#define LA_FLUSH(k, slot) do { if (DBG && dbg_on && ui == 1 && lane == 0) p.dbg[blockIdx.x * 16 + (slot)] = dacc[k]; } while (0)

constexpr int LA_THREADS = 352;
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

// Code size matters here: every role runs its own loop at the same time, and the kernel slowed down by 1.5x whenever the
// loops together outgrew the instruction caches (88 KB of SASS vs 57 KB, profiles/r02_attn_long_probe.txt).  Hence one call
// site per TMA / MMA sequence, rolled k-loops in the issuers, masking of the ragged last tile in TMEM instead of in 64
// registers, and the debug timers compiled into a separate instantiation.
template <bool DBG>
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
  uint64_t* empty = bars + LA_STAGES;               // [STAGES] both MMA issuers -> TMA (2 arrivals)
  uint64_t* bar_q = bars + 2 * LA_STAGES;           // both Q tiles landed
  uint64_t* bar_qfree = bar_q + 1;                  // the unit's last S products retired: Q slots free (2 arrivals)
  uint64_t* bar_s2 = bar_q + 2;                     // [g][half] 64-key half of S_g formed
  uint64_t* bar_p = bar_q + 6;                      // [g][half] that half of P_g written (128 arrivals)
  uint64_t* bar_pv = bar_q + 10;                    // [g][half] P_half.V retired: O_g not in use on its account
  uint64_t* bar_o = bar_q + 14;                     // [g] O_g complete                   (once per unit of the group)
  uint64_t* bar_ofree = bar_q + 16;                 // [g] O_g drained                    (128 arrivals, once per unit of the group)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar_q + 18);
  uint8_t* ostage = tail + 256;                     // [8 softmax warps][32 rows][LA_OPITCH]

  // warp index and TMEM base through a lane-0 shuffle: the compiler then knows they are warp-uniform and keeps the roles'
  // control flow, barrier addresses and UMMA / TMA operands in uniform registers (see the issuer comment below)
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int hd = p.hd, T = p.T, n = p.n_tiles, np = p.n_pairs;
  const int ksteps_s = p.hd16 / 16;
  const int grid = static_cast<int>(gridDim.x);
  const int first_unit = static_cast<int>(blockIdx.x);
  const uint32_t tile_tx = 128u * static_cast<uint32_t>(p.hd16) * 2u;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_q0); tma_prefetch_desc(&map_q1); tma_prefetch_desc(&map_k0);
    tma_prefetch_desc(&map_k1); tma_prefetch_desc(&map_v0); tma_prefetch_desc(&map_v1);
    for (int i = 0; i < LA_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 2); }
    mbar_init(bar_q, 1); mbar_init(bar_qfree, 2);
    for (int i = 0; i < 4; ++i) { mbar_init(&bar_s2[i], 1); mbar_init(&bar_p[i], 128); mbar_init(&bar_pv[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&bar_o[i], 1); mbar_init(&bar_ofree[i], 128); }
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
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);
  // TMEM of group g: [S_g 128 columns | O_g 128 columns] at g * 256; S_g is formed in two 64-key halves, each turned into
  // P in place (32 columns of packed bf16)

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    {
      int stage = 0;
      uint32_t phase = 0;
      auto push = [&](const CUtensorMap* m0, const CUtensorMap* m1, int row) {
        mbar_wait(&empty[stage], phase ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&full[stage], tile_tx);
          uint8_t* dst = ring + stage * LA_TILE;
          tma_load_2d(dst, m0, &full[stage], 0, row);
          tma_load_2d(dst + LA_SUB, m1, &full[stage], 64, row);
        }
        if (++stage == LA_STAGES) { stage = 0; phase ^= 1; }
      };
      int ui = 0;
      for (int u = first_unit; u < p.n_units; u += grid, ++ui) {
        const int item = u / np, pr = u - item * np;
        const int row0 = item * T;
        // K_0, [Q_0 Q_1], V_0, K_1, V_1, ...: one call site (code size); the Q slots are waited for after K_0 is on its way
        for (int i = 0; i < 2 * n; ++i) {
          const int kt = i >> 1;
          push((i & 1) ? &map_v0 : &map_k0, (i & 1) ? &map_v1 : &map_k1, row0 + kt * 128);
          if (i == 0) {
            if (ui > 0) mbar_wait(bar_qfree, (ui - 1) & 1);
            if (elect_one()) {
              mbar_arrive_expect_tx(bar_q, 2 * tile_tx);
              for (int g = 0; g < 2; ++g) {   // a second tile past the head block reads finite neighbour rows, never stored
                tma_load_2d(sQ + g * LA_TILE, &map_q0, bar_q, 0, row0 + (2 * pr + g) * 128);
                tma_load_2d(sQ + g * LA_TILE + LA_SUB, &map_q1, bar_q, 64, row0 + (2 * pr + g) * 128);
              }
            }
          }
        }
      }
    }
  } else if (warp <= 2) {
    // ------------------------------------------------------------------ MMA issuers: warp 1 for group 0, warp 2 for group 1
    // tcgen05.mma of one thread execute in issue order, so the scores of tile kt + 1 issued after the P.V that read the same
    // TMEM columns need no barrier; only the softmax group's tcgen05.st is waited for (bar_p).
    // The WHOLE warp walks the loop and elect.sync picks the lane that issues: with warp-uniform control flow the
    // descriptors, TMEM addresses and barrier addresses stay in uniform registers and UTCHMMA instructions issue back to
    // back.  Under an `if (lane == 0)` around the loop every operand went through R2UR inside an ELECT loop: 20-35
    // dependent instructions = 100-150 cycles per MMA, more than the 32-64 cycles the MMAs of this kernel take, so the
    // issuing threads set the pace (profiles/r02_attn_long_probe.txt).
    {
      const int g = warp - 1;
      const bool dbg_on = p.dbg != nullptr && g == 0;
      long long dacc[7] = {0, 0, 0, 0, 0, 0, 0};
      const int cols_last = p.last_cols;                                  // valid columns of the last key tile (x16)
      const int l0 = cols_last < 64 ? cols_last : 64, l1 = cols_last - l0;  // ... split into its two halves
      const uint32_t id64 = make_idesc_bf16(128, 64), id_l0 = make_idesc_bf16(128, l0), id_l1 = make_idesc_bf16(128, l1 ? l1 : 16);
      const uint32_t idesc_o = make_idesc_bf16(128, p.pv_n) | (1u << 16);   // B (= V) is MN-major
      const uint64_t qdesc = make_smem_desc_sw128(smem_u32(sQ + g * LA_TILE));
      const uint64_t kdesc0 = make_smem_desc_sw128(smem_u32(ring));
      const uint64_t vdesc0 = la_desc_mn(smem_u32(ring), LA_SUB);
      const uint32_t d_s = tmem_base + g * 256, d_o = d_s + 128;
      uint64_t* s2 = bar_s2 + 2 * g; uint64_t* pw = bar_p + 2 * g; uint64_t* pv = bar_pv + 2 * g;
      int stage = 0;
      uint32_t phase = 0;
      int n_p0 = 0, n_p1 = 0;                        // P halves consumed
      int n_mine = 0;                                // units of this group
      int ui = 0;
      auto advance = [&]() { if (++stage == LA_STAGES) { stage = 0; phase ^= 1; } };
      // scores of the keys from key row `row_off` of ring slot st on -> TMEM columns d
      auto issue_s = [&](uint32_t d, int st, int row_off, uint32_t id) {
        const uint64_t kd = kdesc0 + static_cast<uint64_t>(st * (LA_TILE >> 4) + (row_off * 128 >> 4));
#pragma unroll 1
        for (int ks = 0; ks < ksteps_s; ++ks) {
          const uint64_t off = static_cast<uint64_t>((ks >> 2) * (LA_SUB >> 4) + 2 * (ks & 3));
          if (elect_one()) umma_bf16(d, qdesc + off, kd + off, id, ks != 0);
        }
      };
      auto issue_pv = [&](int h, int st, int steps, bool first) {
        const uint64_t vd = vdesc0 + static_cast<uint64_t>(st * (LA_TILE >> 4) + (h * 64 * 128 >> 4));
#pragma unroll 1
        for (int ks = 0; ks < steps; ++ks)
          if (elect_one()) umma_bf16_ts(d_o, d_s + h * 64 + ks * 8, vd + static_cast<uint64_t>(ks * (16 * 128 >> 4)), idesc_o,
                       !(first && ks == 0));
      };
      for (int u = first_unit; u < p.n_units; u += grid, ++ui) {
        const int pr = u % np;
        LA_STAMP(5);
        LA_TIMED_WAIT(0, bar_q, ui & 1);
        if (2 * pr + g >= n) {
          // odd tile count: the last pair has no second query tile; walk the ring so the slot accounting stays in step
          if (elect_one()) mbar_arrive(bar_qfree);
          for (int i = 0; i < 2 * n; ++i) {
            mbar_wait(&full[stage], phase);
            if (elect_one()) mbar_arrive(&empty[stage]);
            advance();
          }
          continue;
        }
        // kt = -1 forms the scores of tile 0; kt >= 0: per half  P_h(kt).V  then the scores S_h(kt + 1) into the same columns
#pragma unroll 1
        for (int kt = -1; kt < n; ++kt) {
          const bool more = kt + 1 < n, last = kt == n - 1, next_last = kt + 1 == n - 1;
          int sv = 0, sk = 0;
          if (kt >= 0) { sv = stage; LA_TIMED_WAIT(1, &full[sv], phase); advance(); }     // V_kt
          if (more)    { sk = stage; LA_TIMED_WAIT(1, &full[sk], phase); advance(); }     // K_{kt+1}
          if (kt == 0 && n_mine > 0) mbar_wait(&bar_ofree[g], (n_mine - 1) & 1);          // the previous unit's O has been drained
#pragma unroll 1
          for (int h = 0; h < 2; ++h) {
            const int cols_h = last ? (h ? l1 : l0) : 64;          // keys of this half in tile kt
            if (kt >= 0 && cols_h > 0) {
              LA_TIMED_WAIT(2, &pw[h], (h ? n_p1 : n_p0) & 1);
              if (h) ++n_p1; else ++n_p0;
              tcgen05_fence_after();
              LA_T0();
              issue_pv(h, sv, cols_h / 16, kt == 0 && h == 0);
              LA_SEG(3);
              if (elect_one()) umma_commit(&pv[h]);
              LA_SEG(5);
            }
            if (more && (!next_last || (h ? l1 : l0) > 0)) {
              LA_T0();
              if (kt < 0) tcgen05_fence_after();
              issue_s(d_s + h * 64, sk, h * 64, next_last ? (h ? id_l1 : id_l0) : id64);
              LA_SEG(4);
              if (elect_one()) umma_commit(&s2[h]);
              LA_SEG(6);
            }
          }
          if (kt >= 0 && elect_one()) umma_commit(&empty[sv]);
          if (more) {
            if (next_last && elect_one()) umma_commit(bar_qfree);   // the unit's last read of the Q tile
            if (elect_one()) umma_commit(&empty[sk]);
          }
        }
        if (elect_one()) umma_commit(&bar_o[g]);
        ++n_mine;
        LA_STAMP(6);
        LA_FLUSH(0, 13); LA_FLUSH(5, 12); LA_FLUSH(6, 11); LA_FLUSH(3, 14); LA_FLUSH(4, 15);
      }
    }
  } else if (warp >= 3) {
    // ------------------------------------------------------------------ softmax groups, thread = query row
    const int g = (warp - 3) >> 2;                    // group g owns query tile 2 * pair + g
    const uint32_t t_s = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + g * 256;
    const uint32_t t_o = t_s + 128;
    uint8_t* patch = ostage + (warp - 3) * (32 * LA_OPITCH);
    uint64_t* s2 = bar_s2 + 2 * g; uint64_t* pw = bar_p + 2 * g; uint64_t* pvb = bar_pv + 2 * g;
    int c_s2a = 0, c_s2b = 0, c_pva = 0, c_pvb = 0, cnt_o = 0;   // uses of the group's barriers so far
    int handed0 = 0, handed1 = 0;                     // P halves handed to the tensor core (each is followed by one bar_pv phase)
    const bool dbg_on = p.dbg != nullptr && warp == 4;   // rows 0-31 of group 0
    long long dacc[4] = {0, 0, 0, 0};
    int ui = 0;
    for (int u = first_unit; u < p.n_units; u += grid, ++ui) {
      const int item = u / np, pr = u - item * np;
      const int qt = 2 * pr + g;
      const int head = item % p.H, bb = item / p.H;
      if (qt >= n) continue;                          // odd tile count: the last pair has no second query tile
      // ---- online softmax over 64-key steps
      LA_STAMP(0);
      float m_ref = -INFINITY;                        // reference maximum of this row, in log2 units (score * scale * log2 e)
      float s4[4] = {0.f, 0.f, 0.f, 0.f};             // running sum of P, four independent chains
      const int unit_h0 = handed0, unit_h1 = handed1; // P halves handed to the tensor core before this unit
      for (int kt = 0; kt < n; ++kt) {
        const int kmax = T - kt * 128;                // keys [0, kmax) of this tile exist
        const int ncols = kt == n - 1 ? p.last_cols : 128;   // columns the tensor core writes
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {                 // not unrolled: the loop body has to stay resident in the instruction cache
          const int cols_h = ncols - h * 64 < 64 ? ncols - h * 64 : 64;
          if (cols_h <= 0) continue;
          LA_TIMED_WAIT(0, &s2[h], (h ? c_s2b : c_s2a) & 1);
          if (h) ++c_s2b; else ++c_s2a;
          tcgen05_fence_after();
          LA_T0();
          const uint32_t t_h = t_s + h * 64;
          if (h * 64 + 64 > kmax) {                   // last tile only (cold): keys past the sequence and the columns the narrow
            const uint32_t ninf = 0xff800000u;        // UMMA did not write become -inf in TMEM, one column at a time
#pragma unroll 1
            for (int c = kmax - h * 64; c < 64; ++c) tmem_st_x1(t_h + c, ninf);
            tmem_st_wait();
          }
          uint32_t v[4][16];
#pragma unroll
          for (int j = 0; j < 4; ++j) tmem_ld_x16(t_h + j * 16, v[j]);
          tmem_ld_wait();
          LA_SEG(1);
          float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
          for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int i = 0; i < 16; ++i) m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(v[j][i]));
          const float m_step = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * p.scale_log2e;
          // O_g must be quiescent before a rescale.  This half's previous P.V retired before these scores were formed (a
          // commit covers everything issued before it): its phase is consumed here without waiting.  The other half's
          // latest P.V may still be running: its phase is consumed if already complete and waited for only when a rescale
          // is due - either way each bar_pv is at most one phase ahead of this thread, so the parity test never aliases.
          if (h) {
            while (c_pvb < handed1) { mbar_wait(&pvb[1], c_pvb & 1); ++c_pvb; }
            if (c_pva < handed0 && mbar_test(&pvb[0], c_pva & 1)) ++c_pva;
          } else {
            while (c_pva < handed0) { mbar_wait(&pvb[0], c_pva & 1); ++c_pva; }
            if (c_pvb < handed1 && mbar_test(&pvb[1], c_pvb & 1)) ++c_pvb;
          }
          LA_SEG(2);
          if (__any_sync(0xffffffffu, m_step > m_ref + 8.f)) {
            while (c_pva < handed0) { mbar_wait(&pvb[0], c_pva & 1); ++c_pva; }
            while (c_pvb < handed1) { mbar_wait(&pvb[1], c_pvb & 1); ++c_pvb; }
            const float m_new = fmaxf(m_ref, m_step);
            const float alpha = la_exp2(m_ref - m_new);            // 0 on the first step (m_ref = -inf)
            if (handed0 + handed1 > unit_h0 + unit_h1) {           // O holds something: scale this warp's rows
              tcgen05_fence_after();
#pragma unroll 1
              for (int c = 0; c < p.pv_n; c += 16) {
                uint32_t o16[16];
                tmem_ld_x16(t_o + c, o16);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) o16[i] = __float_as_uint(__uint_as_float(o16[i]) * alpha);
                tmem_st_x8(t_o + c, o16);
                tmem_st_x8(t_o + c + 8, o16 + 8);
              }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) s4[i] *= alpha;
            m_ref = m_new;
          }
          const float neg_m = -m_ref;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float e[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) e[i] = la_exp2(fmaf(__uint_as_float(v[j][i]), p.scale_log2e, neg_m));
#pragma unroll
            for (int i = 0; i < 16; ++i) s4[i & 3] += e[i];
            uint32_t w[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) w[i] = pack_bf16x2(e[2 * i], e[2 * i + 1]);
            if (j * 16 < cols_h) tmem_st_x8(t_h + j * 8, w);
          }
          tmem_st_wait();
          tcgen05_fence_before();
          mbar_arrive(&pw[h]);
          LA_SEG(3);
          if (h) ++handed1; else ++handed0;
        }
      }
      LA_STAMP(1);
      // ---- epilogue: O_g / rowsum -> bf16 -> shared-memory transpose -> global
      const float sum = (s4[0] + s4[1]) + (s4[2] + s4[3]);
      LA_STAMP(2);
      mbar_wait(&bar_o[g], cnt_o & 1);
      ++cnt_o;
      while (c_pva < handed0) { mbar_wait(&pvb[0], c_pva & 1); ++c_pva; }   // complete by now: keep the phase counts in step
      while (c_pvb < handed1) { mbar_wait(&pvb[1], c_pvb & 1); ++c_pvb; }
      LA_STAMP(3);
      tcgen05_fence_after();
      const float inv = sum > 0.f ? 1.f / sum : 0.f;
      const int q_warp0 = qt * 128 + (warp & 3) * 32;
      __nv_bfloat16* obase = p.o + (static_cast<long long>(bb) * T + q_warp0) * p.ldo + head * hd;
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
      mbar_arrive(&bar_ofree[g]);
      LA_STAMP(4);
      LA_FLUSH(0, 9); LA_FLUSH(1, 7); LA_FLUSH(2, 8); LA_FLUSH(3, 10);
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
  p.dbg = reinterpret_cast<long long*>(getenv("CGPT_ATTN_DBG") ? strtoull(getenv("CGPT_ATTN_DBG"), nullptr, 0) : 0ull);
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
    CGPT_CHECK_CUDA(cudaFuncSetAttribute(attn_long_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CGPT_CHECK_CUDA(cudaFuncSetAttribute(attn_long_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  int grid = sms;
  if (grid > p.n_units) grid = p.n_units;
  if (p.dbg) attn_long_kernel<true><<<grid, LA_THREADS, smem, stream>>>(mq0, mq1, mk0, mk1, mv0, mv1, p);
  else       attn_long_kernel<false><<<grid, LA_THREADS, smem, stream>>>(mq0, mq1, mk0, mk1, mv0, mv1, p);
  CGPT_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace cgpt
