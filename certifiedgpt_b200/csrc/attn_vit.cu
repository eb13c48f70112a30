// Pipelined EVA ViT self-attention on the 5th-gen tensor cores (tcgen05 + TMEM + TMA)       (eva_vit.py:123-153)
//   T = 257 = 1 cls + 256 patches, 16 heads x 88, non-causal; also every non-causal shape with <= 256 (+1) keys.
//
// Inputs are HEAD-MAJOR: q, k, v each [B][H][T][hd] bf16, written in that layout by the fused-QKV GEMM epilogue
// (cgpt_gemm_epilogue.hm_*), so one (sample, head) item is ONE contiguous block per operand (257 x 176 B) instead of 257
// 176-byte strips of an 8448-byte row: every TMA box is a dense stream (the strip layout read 2.2x the algorithmic
// bytes from HBM, profiles/r01_attn_umma_final_ncu.txt) and the columns hd..roundup16(hd) come back as TMA zero fill.
//
// One persistent CTA per SM walks its items as a software pipeline over (item, 128-query tile) units:
//   warp 0     TMA producer : Q tile(s), K, V of item i+1 into the slots item i has released (+ L2 prefetch of i+2)
//   warp 1     MMA issuer   : one thread, over the two TMEM regions:
//                               S_t = Q_t K^T  (UMMA 128 x nk x 16, fp32 in TMEM region t, 256 columns)
//                               O_t = P_t V    (A operand = P_t READ FROM TMEM, B = V in place, MN-major)
//   warps 4-7  softmax of q-tile 0, warps 8-11 softmax of q-tile 1 (thread = query row): row max, exp2, and the
//              unnormalised P as packed bf16 written IN PLACE over the first 128 columns of its own S region with
//              tcgen05.st (no shared-memory P tile, no proxy fence); O_t lands in columns 128.. of the same region.
//   warps 2-3  the cls QUERY row on CUDA cores (scores from the softmax threads, P.V as a GEMV out of the V tile)
// S_t of item i+1 is issued as soon as O_t of item i has been drained, and the softmax threads compute the cls scores
// of item i+1 while P_t.V of item i runs, so the tensor core, the softmax groups (MUFU-bound) and the loads of the next
// item overlap; the cls KEY is folded in on CUDA cores (one extra score per row + a rank-1 update of O) so every
// tensor-core tile is exactly 128 x 256.  O leaves through a per-warp shared-memory transpose: every global store
// instruction writes 8 rows x 64 contiguous bytes instead of 32 rows x 16.
#include <stdlib.h>
#include "common.cuh"
#include "ops.h"

namespace cgpt {

struct VitAttnParams {
  const __nv_bfloat16* q; const __nv_bfloat16* k; const __nv_bfloat16* v;   // head-major [B][H][T][hd]
  __nv_bfloat16* o; long long ldo;                                           // row-major [B*T, >= H*hd]
  int H, hd, hd16, T, E;        // T includes the E extra (cls) row, which is row 0 of every head block
  float scale_log2e;
  int n_qtiles, nk_pad, pv_n;   // 128-row query tiles; main keys padded to a multiple of 16; UMMA N of P.V
  int n_items;                  // B * H
  int prefetch;                 // != 0: L2 prefetch two items ahead
  long long* dbg;               // optional [grid][16] cycle stamps (CGPT_ATTN_DBG)
};

constexpr int VA_THREADS = 384;
constexpr int VSUB = 16384;   // one 128-row x 64-col bf16 sub-tile (128-byte rows, SWIZZLE_128B)

__device__ __forceinline__ uint64_t va_desc_kmajor(uint32_t addr) { return make_smem_desc_sw128(addr); }
// MN-major operand, 128B swizzle: LBO = byte stride between 64-element MN atoms, SBO = 8-row K-group stride
__device__ __forceinline__ uint64_t va_desc_mnmajor(uint32_t addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>(1024u >> 4) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}
__device__ __forceinline__ float va_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(VA_THREADS, 1)
attn_vit_kernel(const __grid_constant__ CUtensorMap map_q0, const __grid_constant__ CUtensorMap map_q1,
                const __grid_constant__ CUtensorMap map_k0, const __grid_constant__ CUtensorMap map_k1,
                const __grid_constant__ CUtensorMap map_v0, const __grid_constant__ CUtensorMap map_v1,
                VitAttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int kv_sub = p.nk_pad * 128;
  uint8_t* sQ = smem;                                  // n_qtiles x [sub0 | sub1]
  uint8_t* sK = sQ + p.n_qtiles * 2 * VSUB;            // [sub0 | sub1], sub stride kv_sub
  uint8_t* sV = sK + 2 * kv_sub;                       // [sub0 | sub1]
  uint8_t* tail = sV + 2 * kv_sub;
  // every barrier completes exactly once per item of this CTA: item `it` waits with parity it & 1
  // (bar_xr: once per use of an x-buffer, parity (it >> 1) & 1)
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail);
  uint64_t* bar_q = bars;          // [2] TMA: Q tile t landed
  uint64_t* bar_k = bars + 2;      //     TMA: K landed
  uint64_t* bar_v = bars + 3;      //     TMA: V landed
  uint64_t* bar_s = bars + 4;      // [2] S_t in TMEM (tcgen05.commit) = Q tile t / K consumed by the tensor core
  uint64_t* bar_p = bars + 6;      // [2] P_t in TMEM (128 softmax threads)
  uint64_t* bar_o = bars + 8;      // [2] O_t in TMEM (tcgen05.commit) = V consumed by the tensor core
  uint64_t* bar_free = bars + 10;  // [2] region t drained by its softmax group (128 arrivals)
  uint64_t* bar_qkr = bars + 12;   //     softmax threads no longer read the Q / K tiles (cls scores)
  uint64_t* bar_vf = bars + 13;    //     cls-query warps no longer read the V tile (64 arrivals)
  uint64_t* bar_x = bars + 14;     //     cls-query scores written (one arrival per softmax thread)
  uint64_t* bar_xr = bars + 15;    // [2] cls key / value / query of an item staged in x-buffer [it & 1] (32 arrivals)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 17);
  constexpr int XBUF = 3 * 128 + 320;                  // [xk 128 | xv 128 | xq 128 | xs 320] fp32
  float* xbase = reinterpret_cast<float*>(tail + 256);
  float* xo = xbase + 2 * XBUF;                        // cls-query partial outputs [5][96+]
  constexpr int OPITCH = 80;                           // bytes per staged row: 32 bf16 + 16 B pad (conflict-free 16-byte writes)
  uint8_t* ostage = reinterpret_cast<uint8_t*>(xo + 5 * 96 + 32);   // [8 softmax warps][32 rows][OPITCH]

  // warp index and TMEM base through a lane-0 shuffle: known warp-uniform to the compiler (see the MMA issuer)
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
#define VA_STAMP(slot) do { if (p.dbg) p.dbg[blockIdx.x * 16 + (slot)] = clock64(); } while (0)
  const int hd = p.hd, E = p.E, T = p.T, nqt = p.n_qtiles;
  const int Tk_main = T - E, Tq_main = T - E;
  const int ksteps_s = p.hd16 / 16;
  const int ksteps_o = p.nk_pad / 16;
  const int grid = static_cast<int>(gridDim.x);
  const int n_soft = 128 * nqt;
  const int first_item = static_cast<int>(blockIdx.x);
  const int n_mine = first_item < p.n_items ? (p.n_items - first_item + grid - 1) / grid : 0;
  const uint32_t q_tx = 128u * static_cast<uint32_t>(p.hd16) * 2u;
  const uint32_t kv_tx = static_cast<uint32_t>(p.nk_pad) * static_cast<uint32_t>(p.hd16) * 2u;

  auto load_q = [&](int item, int t) {
    const int row = item * T + E + t * 128;
    mbar_arrive_expect_tx(&bar_q[t], q_tx);
    tma_load_2d(sQ + t * 2 * VSUB, &map_q0, &bar_q[t], 0, row);
    tma_load_2d(sQ + t * 2 * VSUB + VSUB, &map_q1, &bar_q[t], 64, row);
  };
  auto load_k = [&](int item) {
    const int row = item * T + E;
    mbar_arrive_expect_tx(bar_k, kv_tx);
    tma_load_2d(sK, &map_k0, bar_k, 0, row);
    tma_load_2d(sK + kv_sub, &map_k1, bar_k, 64, row);
  };
  auto load_v = [&](int item) {
    const int row = item * T + E;
    mbar_arrive_expect_tx(bar_v, kv_tx);
    tma_load_2d(sV, &map_v0, bar_v, 0, row);
    tma_load_2d(sV + kv_sub, &map_v1, bar_v, 64, row);
  };
  auto prefetch_item = [&](int item) {
    if (!p.prefetch || item >= p.n_items) return;
    const int row = item * T + E;
    for (int t = 0; t < nqt; ++t) {
      tma_prefetch_l2_2d(&map_q0, 0, row + t * 128);
      tma_prefetch_l2_2d(&map_q1, 64, row + t * 128);
    }
    tma_prefetch_l2_2d(&map_k0, 0, row);
    tma_prefetch_l2_2d(&map_k1, 64, row);
    tma_prefetch_l2_2d(&map_v0, 0, row);
    tma_prefetch_l2_2d(&map_v1, 64, row);
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_q0); tma_prefetch_desc(&map_q1); tma_prefetch_desc(&map_k0);
    tma_prefetch_desc(&map_k1); tma_prefetch_desc(&map_v0); tma_prefetch_desc(&map_v1);
    mbar_init(&bar_q[0], 1); mbar_init(&bar_q[1], 1);
    mbar_init(bar_k, 1); mbar_init(bar_v, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_s[i], 1); mbar_init(&bar_p[i], 128); mbar_init(&bar_o[i], 1); mbar_init(&bar_free[i], 128);
    }
    mbar_init(bar_qkr, n_soft);
    mbar_init(bar_vf, 64);
    mbar_init(bar_x, n_soft);
    mbar_init(&bar_xr[0], 32);
    mbar_init(&bar_xr[1], 32);
    fence_barrier_init();
    fence_proxy_async();
    if (n_mine > 0) {
      load_k(first_item);
      for (int t = 0; t < nqt; ++t) load_q(first_item, t);
      load_v(first_item);
      prefetch_item(first_item + grid);
    }
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_ptr, 0);

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int it = 0;
      for (int item = first_item; item < p.n_items; item += grid, ++it) {
        const int next = item + grid;
        if (next >= p.n_items) break;
        const uint32_t ph = it & 1;
        prefetch_item(item + 2 * grid);
        // Q tile 0 of the next item: the tensor core (S_0) and the softmax threads (cls scores) are done with it
        mbar_wait(bar_qkr, ph);
        mbar_wait(&bar_s[0], ph);
        load_q(next, 0);
        if (nqt == 2) mbar_wait(&bar_s[1], ph);
        load_k(next);
        if (nqt == 2) load_q(next, 1);
        // V: both P.V products retired, cls-query GEMV done
        mbar_wait(&bar_o[0], ph);
        if (nqt == 2) mbar_wait(&bar_o[1], ph);
        if (E) mbar_wait(bar_vf, ph);
        load_v(next);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    // The whole warp walks the loop and elect.sync picks the issuing lane: with warp-uniform control flow the UMMA
    // operands stay in uniform registers and UTCHMMA issues back to back.  Under `if (lane == 0)` every operand went
    // through R2UR inside an ELECT loop, 100-170 cycles per MMA (profiles/r02_attn_long_probe.txt) - three times what
    // the 48-cycle P.V MMAs take.
    if (n_mine > 0) {
      const uint32_t idesc_s = make_idesc_bf16(128, p.nk_pad);
      const uint32_t idesc_o = make_idesc_bf16(128, p.pv_n) | (1u << 16);   // B (= V) is MN-major
      auto issue_s = [&](int t) {
        const uint32_t d_tmem = tmem_base + t * 256;
        for (int ks = 0; ks < ksteps_s; ++ks) {
          const uint64_t ad = va_desc_kmajor(smem_u32(sQ + t * 2 * VSUB + (ks >> 2) * VSUB)) + 2 * (ks & 3);
          const uint64_t bd = va_desc_kmajor(smem_u32(sK + (ks >> 2) * kv_sub)) + 2 * (ks & 3);
          if (elect_one()) umma_bf16(d_tmem, ad, bd, idesc_s, ks != 0);
        }
        if (elect_one()) umma_commit(&bar_s[t]);
      };
      auto issue_pv = [&](int t) {
        const uint32_t d_tmem = tmem_base + t * 256 + 128;
        const uint32_t a_tmem = tmem_base + t * 256;       // P_t: 8 columns (16 bf16) per K step
        for (int ks = 0; ks < ksteps_o; ++ks) {
          const uint64_t bd = va_desc_mnmajor(smem_u32(sV + ks * 16 * 128), static_cast<uint32_t>(kv_sub));
          if (elect_one()) umma_bf16_ts(d_tmem, a_tmem + ks * 8, bd, idesc_o, ks != 0);
        }
        if (elect_one()) umma_commit(&bar_o[t]);
      };
      mbar_wait(bar_k, 0);
      for (int t = 0; t < nqt; ++t) {
        mbar_wait(&bar_q[t], 0);
        tcgen05_fence_after();
        issue_s(t);
      }
      if (lane == 0) VA_STAMP(2);
      // both softmax groups run in lockstep (same item, same phase), so the natural order is fixed:
      // P.V of both tiles as their P arrives, then S of the next item for each tile as its region drains
      for (int it = 0; it < n_mine; ++it) {
        const uint32_t ph = it & 1;
        mbar_wait(bar_v, ph);
        for (int t = 0; t < nqt; ++t) {
          mbar_wait(&bar_p[t], ph);
          tcgen05_fence_after();
          issue_pv(t);
        }
        if (it + 1 < n_mine) {
          mbar_wait(bar_k, ph ^ 1);
          for (int t = 0; t < nqt; ++t) {
            mbar_wait(&bar_free[t], ph);       // region t drained by its softmax group
            mbar_wait(&bar_q[t], ph ^ 1);      // Q tile t of the next item landed
            tcgen05_fence_after();
            issue_s(t);
          }
        }
      }
      if (lane == 0) VA_STAMP(3);
    }
  } else if (warp == 2 || warp == 3) {
    // ------------------------------------------------------------------ cls query row, CUDA cores
    if (E) {
      const int t64 = threadIdx.x - 64;   // 0..63
      int it = 0;
      for (int item = first_item; item < p.n_items; item += grid, ++it) {
        const uint32_t ph = it & 1;
        const int h = item % p.H, b = item / p.H;
        float* xb = xbase + (it & 1) * XBUF;
        float* xk = xb; float* xv = xb + 128; float* xq = xb + 256; float* xs = xb + 384;
        if (warp == 3) {
          // stage this item's cls key / value / query as fp32; the buffer was last used two items ago, whose
          // epilogue (the rank-1 cls term reads xv) runs AFTER the softmax threads' cls scores of item it - 1
          if (it >= 2)
            for (int t = 0; t < nqt; ++t) mbar_wait(&bar_free[t], ph);
          const long long row0 = static_cast<long long>(item) * T * hd;
          if (lane < 16) {
            const bool ok = lane * 8 < hd;
            uint4 a = make_uint4(0, 0, 0, 0), bq = a, c = a;
            if (ok) {
              a = *reinterpret_cast<const uint4*>(p.k + row0 + lane * 8);
              bq = *reinterpret_cast<const uint4*>(p.v + row0 + lane * 8);
              c = *reinterpret_cast<const uint4*>(p.q + row0 + lane * 8);
            }
            const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {bq.x, bq.y, bq.z, bq.w}, cw[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              xk[lane * 8 + 2 * i] = bf16_lo(aw[i]); xk[lane * 8 + 2 * i + 1] = bf16_hi(aw[i]);
              xv[lane * 8 + 2 * i] = bf16_lo(bw[i]); xv[lane * 8 + 2 * i + 1] = bf16_hi(bw[i]);
              xq[lane * 8 + 2 * i] = bf16_lo(cw[i]) * p.scale_log2e; xq[lane * 8 + 2 * i + 1] = bf16_hi(cw[i]) * p.scale_log2e;
            }
          }
          mbar_arrive(&bar_xr[it & 1]);
        }
        mbar_wait(&bar_xr[it & 1], (it >> 1) & 1);
        if (t64 == 0) {
          float acc = 0.f;
          for (int d = 0; d < hd; ++d) acc = fmaf(xq[d], xk[d], acc);
          xs[0] = acc;                       // cls query . cls key
        }
        mbar_wait(bar_x, ph);
        asm volatile("bar.sync 2, 64;" ::: "memory");
        if (warp == 2) {
          float mx = -INFINITY;
          for (int j = lane; j < T; j += 32) mx = fmaxf(mx, xs[j]);
          mx = warp_max(mx);
          float sum = 0.f;
          for (int j = lane; j < T; j += 32) {
            const float e = va_exp2(xs[j] - mx);
            xs[j] = e;
            sum += e;
          }
          sum = warp_sum(sum);
          if (lane == 0) xs[T] = 1.f / sum;
        }
        mbar_wait(bar_v, ph);
        asm volatile("bar.sync 2, 64;" ::: "memory");
        // (hd / 8) column chunks x 5 key partitions <= 64 threads (hd <= 96)
        const int nch = hd >> 3;
        const int c = t64 % nch, part = t64 / nch;
        if (part < 5) {
          float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          const int per = (Tk_main + 4) / 5;
          const int j0 = part * per, j1 = min(Tk_main, j0 + per);
          const uint8_t* vsub = sV + (c >> 3) * kv_sub;
          for (int j = j0; j < j1; ++j) {
            const uint4 vv = *reinterpret_cast<const uint4*>(vsub + j * 128 + (((c & 7) ^ (j & 7)) << 4));
            const float pj = xs[1 + j];
            acc[0] = fmaf(pj, bf16_lo(vv.x), acc[0]); acc[1] = fmaf(pj, bf16_hi(vv.x), acc[1]);
            acc[2] = fmaf(pj, bf16_lo(vv.y), acc[2]); acc[3] = fmaf(pj, bf16_hi(vv.y), acc[3]);
            acc[4] = fmaf(pj, bf16_lo(vv.z), acc[4]); acc[5] = fmaf(pj, bf16_hi(vv.z), acc[5]);
            acc[6] = fmaf(pj, bf16_lo(vv.w), acc[6]); acc[7] = fmaf(pj, bf16_hi(vv.w), acc[7]);
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) xo[part * 96 + c * 8 + i] = acc[i];
        }
        mbar_arrive(bar_vf);                 // this thread no longer reads the V tile
        asm volatile("bar.sync 2, 64;" ::: "memory");
        if (t64 < nch) {
          const float inv = xs[T];
          float f[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int d = t64 * 8 + i;
            f[i] = (xo[d] + xo[96 + d] + xo[192 + d] + xo[288 + d] + xo[384 + d] + xs[0] * xv[d]) * inv;
          }
          *reinterpret_cast<uint4*>(p.o + static_cast<long long>(b) * T * p.ldo + h * hd + t64 * 8) =
              make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
        }
        asm volatile("bar.sync 2, 64;" ::: "memory");   // xo is reused by the next item
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax + epilogue, thread = query row
    const int t = (warp - 4) >> 2;
    if (t < nqt) {
      const int r = (warp & 3) * 32 + lane;
      const int q_main = t * 128 + r;                  // index among the main query rows (= main key index for the cls scores)
      const bool row_ok = q_main < Tq_main;
      const uint32_t t_lane = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + t * 256;
      const int kmax = Tk_main;                        // main keys [0, kmax) visible (padded keys masked)
      uint8_t* patch = ostage + (warp - 4) * (32 * OPITCH);
      // Scores against the cls key (this row, returned) and of the cls query against main key q_main (thread <-> key,
      // into xs), read from the TMA-loaded, swizzled Q / K tiles in shared memory.  All the 16-byte loads of a row are
      // issued before the first multiply (hd <= 96: at most 12 chunks).
      // This row against the cls key (returned) and the cls query against main key q_main (into xs): the item's share of
      // CUDA-core dot products, computed while P_t.V of the previous item is in flight.
      uint4 tile[12];
      auto dot = [&](const float* x, int nch) {
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < 12; ++c) {
          if (c < nch) {
            const float4 x0 = *reinterpret_cast<const float4*>(x + c * 8), x1 = *reinterpret_cast<const float4*>(x + c * 8 + 4);
            acc += bf16_lo(tile[c].x) * x0.x + bf16_hi(tile[c].x) * x0.y + bf16_lo(tile[c].y) * x0.z + bf16_hi(tile[c].y) * x0.w +
                   bf16_lo(tile[c].z) * x1.x + bf16_hi(tile[c].z) * x1.y + bf16_lo(tile[c].w) * x1.z + bf16_hi(tile[c].w) * x1.w;
          }
        }
        return acc;
      };
      auto cls_key_score = [&](int it) -> float {
        const float* xk = xbase + (it & 1) * XBUF;
        mbar_wait(&bar_xr[it & 1], (it >> 1) & 1);
        mbar_wait(&bar_q[t], it & 1);
        const int nch = hd >> 3;
        const uint8_t* qrow = sQ + t * 2 * VSUB + r * 128;
#pragma unroll
        for (int c = 0; c < 12; ++c)
          if (c < nch) tile[c] = *reinterpret_cast<const uint4*>(qrow + (c >> 3) * VSUB + (((c & 7) ^ (r & 7)) << 4));
        const float a1 = dot(xk, nch);
        return row_ok ? a1 : -INFINITY;
      };
      auto cls_query_score = [&](int it) {
        float* xb = xbase + (it & 1) * XBUF;
        mbar_wait(bar_k, it & 1);
        if (q_main < Tk_main) {
          const int nch = hd >> 3;
          const uint8_t* krow = sK + q_main * 128;        // key index == q_main (requires q_main < nk_pad)
#pragma unroll
          for (int c = 0; c < 12; ++c)
            if (c < nch) tile[c] = *reinterpret_cast<const uint4*>(krow + (c >> 3) * kv_sub + (((c & 7) ^ (q_main & 7)) << 4));
          xb[384 + 1 + q_main] = dot(xb + 256, nch);
        }
        mbar_arrive(bar_x);
      };
      float s_x = -INFINITY;
      if (n_mine > 0) {
        if (E) { s_x = cls_key_score(0); cls_query_score(0); }
        mbar_arrive(bar_qkr);              // this thread no longer reads the Q / K tiles of item 0
      }
      int it = 0;
      for (int item = first_item; item < p.n_items; item += grid, ++it) {
        const uint32_t ph = it & 1;
        const int h = item % p.H, b = item / p.H;
        const float* xv = xbase + (it & 1) * XBUF + 128;
        const bool stamp = threadIdx.x == 128 && it == (n_mine > 2 ? 1 : 0);   // a steady-state item
        if (stamp) VA_STAMP(5);
        mbar_wait(&bar_s[t], ph);
        if (stamp) VA_STAMP(7);
        tcgen05_fence_after();
        // pass 1: row max of the raw scores (mask: padded keys).  16-column TMEM loads, the next one in flight
        // while the current is reduced; full chunks skip the mask.
        float mx = s_x;
        {
          uint32_t va[16], vb[16];
          tmem_ld_x16(t_lane, va);
#pragma unroll 1
          for (int c = 0; c < p.nk_pad; c += 32) {
            tmem_ld_wait();
            if (c + 16 < p.nk_pad) tmem_ld_x16(t_lane + c + 16, vb);
            if (c + 16 <= kmax) {
#pragma unroll
              for (int i = 0; i < 16; ++i) mx = fmaxf(mx, __uint_as_float(va[i]));
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) if (c + i < kmax) mx = fmaxf(mx, __uint_as_float(va[i]));
            }
            if (c + 16 < p.nk_pad) {
              tmem_ld_wait();
              if (c + 32 < p.nk_pad) tmem_ld_x16(t_lane + c + 32, va);
              if (c + 32 <= kmax) {
#pragma unroll
                for (int i = 0; i < 16; ++i) mx = fmaxf(mx, __uint_as_float(vb[i]));
              } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) if (c + 16 + i < kmax) mx = fmaxf(mx, __uint_as_float(vb[i]));
              }
            }
          }
        }
        if (mx == -INFINITY) mx = 0.f;
        if (stamp) VA_STAMP(8);
        const float neg_ms = -mx * p.scale_log2e;
        // pass 2: P = exp2(s * scale - max * scale) as packed bf16, IN PLACE over the S region: the 16 keys of
        // chunk c go to columns [c / 2, c / 2 + 8), always behind the columns still to be read
        float sum = 0.f;
        auto emit = [&](const uint32_t* v, int c) {
          float e[16];
          if (c + 16 <= kmax) {
#pragma unroll
            for (int i = 0; i < 16; ++i) e[i] = va_exp2(fmaf(__uint_as_float(v[i]), p.scale_log2e, neg_ms));
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              e[i] = (c + i < kmax) ? va_exp2(fmaf(__uint_as_float(v[i]), p.scale_log2e, neg_ms)) : 0.f;
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) sum += e[i];
          uint32_t w[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) w[i] = pack_bf16x2(e[2 * i], e[2 * i + 1]);
          tmem_st_x8(t_lane + (c >> 1), w);
        };
        {
          uint32_t va[16], vb[16];
          tmem_ld_x16(t_lane, va);
#pragma unroll 1
          for (int c = 0; c < p.nk_pad; c += 32) {
            tmem_ld_wait();
            if (c + 16 < p.nk_pad) tmem_ld_x16(t_lane + c + 16, vb);
            emit(va, c);
            if (c + 16 < p.nk_pad) {
              tmem_ld_wait();
              if (c + 32 < p.nk_pad) tmem_ld_x16(t_lane + c + 32, va);
              emit(vb, c + 16);
            }
          }
        }
        float p_x = 0.f;
        if (E && row_ok) {
          p_x = va_exp2(fmaf(s_x, p.scale_log2e, neg_ms));
          sum += p_x;
        }
        tmem_st_wait();             // this thread's P stores have landed in TMEM
        tcgen05_fence_before();     // ... and are ordered before the P.V product the MMA thread issues after the barrier
        mbar_arrive(&bar_p[t]);
        if (stamp) VA_STAMP(9);
        // while P_t.V runs: the cls scores of the NEXT item (its Q tile and K landed long ago).  Measured: with the
        // query half moved in front of the S wait instead, the item took 15.7k cycles against 14.9k - the S products
        // keep the shared-memory port busy (~96 B/clk of 128), P.V only half of it.
        if (item + grid < p.n_items) {
          if (E) { s_x = cls_key_score(it + 1); cls_query_score(it + 1); }
          mbar_arrive(bar_qkr);            // this thread no longer reads the Q / K tiles of item it + 1
        }
        if (stamp) VA_STAMP(12);
        // ---- epilogue: O_t / rowsum (+ the cls key's rank-1 term) -> bf16 -> shared-memory transpose -> global
        mbar_wait(&bar_o[t], ph);
        if (stamp) VA_STAMP(10);
        tcgen05_fence_after();
        const float inv = sum > 0.f ? 1.f / sum : 0.f;
        const uint32_t o_addr = t_lane + 128;
        // row served by this lane on the way out: (lane >> 2) + 8 * k of the warp's 32 rows, 16-byte segment lane & 3
        const int q_warp0 = t * 128 + (warp & 3) * 32;
        __nv_bfloat16* obase = p.o + (static_cast<long long>(b) * T + E + q_warp0) * p.ldo + h * hd;
#pragma unroll 1
        for (int c = 0; c < hd; c += 32) {
          uint32_t v[32];
          tmem_ld_x32(o_addr + c, v);
          tmem_ld_wait();
          uint32_t w[16];
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            float f0 = __uint_as_float(v[i]), f1 = __uint_as_float(v[i + 1]), f2 = __uint_as_float(v[i + 2]),
                  f3 = __uint_as_float(v[i + 3]);
            if (E) {
              const float4 x4 = *reinterpret_cast<const float4*>(xv + ((c + i) & 127));   // one broadcast 16-byte load
              f0 += p_x * x4.x; f1 += p_x * x4.y; f2 += p_x * x4.z; f3 += p_x * x4.w;
            }
            w[i >> 1] = pack_bf16x2(f0 * inv, f1 * inv);
            w[(i >> 1) + 1] = pack_bf16x2(f2 * inv, f3 * inv);
          }
          uint4* mine = reinterpret_cast<uint4*>(patch + lane * OPITCH);
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
              if (q_warp0 + rr < Tq_main)
                *reinterpret_cast<uint4*>(obase + static_cast<long long>(rr) * p.ldo + c + seg * 8) =
                    *reinterpret_cast<const uint4*>(patch + rr * OPITCH + seg * 16);
            }
          }
          __syncwarp();
        }
        tcgen05_fence_before();     // this thread's TMEM reads of O_t are done before the next S_t may overwrite the region
        mbar_arrive(&bar_free[t]);
        if (stamp) VA_STAMP(11);
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
#undef VA_STAMP
}

// ------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn3)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn3 va_encode_fn() {
  static EncodeTiledFn3 fn = nullptr;
  if (fn) return fn;
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn3>(ptr);
  return fn;
}
// [rows, cols] bf16, dense rows of `cols` elements; boxes past column `cols` or row `rows` are zero-filled
static int va_make_map(CUtensorMap* map, const void* base, long long rows, int cols, int box_cols, int box_rows) {
  EncodeTiledFn3 fn = va_encode_fn();
  CGPT_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled driver entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CGPT_REQUIRE(r == CUDA_SUCCESS, "attention_vit: cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%d box=%dx%d", (int)r,
               rows, cols, box_cols, box_rows);
  return 0;
}

// 1 = this head-major problem is served by the pipelined tcgen05 kernel
int attn_vit_supported(const cgpt_attn_args* a) {
  if (!a->head_major || a->causal || a->P != 0 || a->decode_kernel != 0) return 0;
  if (a->head_dim <= 64 || a->head_dim > 128 || (a->head_dim & 7)) return 0;
  if (a->Tq != a->Tk || a->q_rows_per_batch != a->Tq || a->kv_rows_per_batch != a->Tk) return 0;
  if (a->B * (long long)a->H * a->Tk > 0x7fffffffLL) return 0;
  const int E = a->Tk > 256 ? 1 : 0;
  if (a->Tk - E > 256 || a->Tk - E < 1) return 0;
  if (E && a->head_dim > 96) return 0;    // the cls-query GEMV maps (hd / 8) x 5 partitions onto 64 threads
  if ((reinterpret_cast<uintptr_t>(a->q) | reinterpret_cast<uintptr_t>(a->k) | reinterpret_cast<uintptr_t>(a->v) |
       reinterpret_cast<uintptr_t>(a->o)) & 15)
    return 0;
  return 1;
}

int attention_vit(const cgpt_attn_args* a, cudaStream_t stream) {
  CGPT_REQUIRE(attn_vit_supported(a), "attention_vit: unsupported head-major problem (Tq=%d Tk=%d hd=%d causal=%d)", a->Tq,
               a->Tk, a->head_dim, a->causal);
  VitAttnParams p;
  p.q = (const __nv_bfloat16*)a->q; p.k = (const __nv_bfloat16*)a->k; p.v = (const __nv_bfloat16*)a->v;
  p.o = (__nv_bfloat16*)a->o; p.ldo = a->ldo;
  p.H = a->H; p.hd = a->head_dim; p.hd16 = (a->head_dim + 15) & ~15; p.T = a->Tk;
  p.E = a->Tk > 256 ? 1 : 0;
  p.scale_log2e = a->scale * 1.4426950408889634f;
  const int t_main = a->Tk - p.E;
  p.n_qtiles = (t_main + 127) / 128;
  p.nk_pad = (t_main + 15) / 16 * 16;
  static const bool pv_wide = getenv("CGPT_ATTN_PV_N128") != nullptr;   // A/B: P.V with UMMA N = 128 instead of roundup16(hd)
  p.pv_n = pv_wide ? 128 : p.hd16;
  p.n_items = a->B * a->H;
  p.prefetch = getenv("CGPT_ATTN_NO_PREFETCH") ? 0 : 1;
  p.dbg = reinterpret_cast<long long*>(getenv("CGPT_ATTN_DBG") ? strtoull(getenv("CGPT_ATTN_DBG"), nullptr, 0) : 0ull);
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const int kv_sub = p.nk_pad * 128;
  const int smem = p.n_qtiles * 2 * VSUB + 4 * kv_sub + 256 + (2 * (3 * 128 + 320) + 5 * 96 + 32) * 4 +
                   8 * 32 * 80 /* epilogue transpose patches */ + 1024;
  CGPT_REQUIRE(smem <= 227 * 1024, "attention_vit: shared memory %d too large", smem);
  const long long rows = (long long)a->B * a->H * a->Tk;
  CUtensorMap mq0, mq1, mk0, mk1, mv0, mv1;
  const int c1 = p.hd16 - 64;
  if (int rc = va_make_map(&mq0, a->q, rows, a->head_dim, 64, 128)) return rc;
  if (int rc = va_make_map(&mq1, a->q, rows, a->head_dim, c1, 128)) return rc;
  if (int rc = va_make_map(&mk0, a->k, rows, a->head_dim, 64, p.nk_pad)) return rc;
  if (int rc = va_make_map(&mk1, a->k, rows, a->head_dim, c1, p.nk_pad)) return rc;
  if (int rc = va_make_map(&mv0, a->v, rows, a->head_dim, 64, p.nk_pad)) return rc;
  if (int rc = va_make_map(&mv1, a->v, rows, a->head_dim, c1, p.nk_pad)) return rc;
  static int configured_smem = 0;
  if (smem > configured_smem) {
    CGPT_CHECK_CUDA(cudaFuncSetAttribute(attn_vit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured_smem = smem;
  }
  int grid = sms;                      // persistent: one CTA per SM walks items blockIdx.x, blockIdx.x + grid, ...
  if (grid > p.n_items) grid = p.n_items;
  attn_vit_kernel<<<grid, VA_THREADS, smem, stream>>>(mq0, mq1, mk0, mk1, mv0, mv1, p);
  CGPT_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace cgpt
