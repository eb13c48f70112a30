// One-shot attention on the 5th-gen tensor cores (tcgen05 + TMEM + TMA) for the two shapes that
// dominate the non-GEMM time of the smoothing path:
//   - EVA ViT self-attention, T = 257 = 1 cls + 256 patches, 16 heads x 88       (eva_vit.py:133-150)
//   - Llama prefill over [prefix | image | question] rows, <= 256 keys, causal     (HF LlamaAttention)
//
// One CTA per (batch, head).  All keys fit one UMMA N (<= 256), so there is no online-softmax loop:
//   TMA   : Q tile(s) [128 x hd], K [nk x hd], V [nk x hd] -> 128B-swizzled shared memory
//   UMMA  : S = Q K^T (M=128, N=nk, K=hd in 16-steps)            -> TMEM, fp32
//   warps : thread-per-row softmax straight out of TMEM (two passes: max, then exp/sum),
//           unnormalised P as bf16 into shared memory (K-major, 128B swizzle) over the dead K/Q tiles
//   UMMA  : O = P V  (M=128, N=128, K=nk; V is the MN-major B operand, read in place)  -> TMEM
//   warps : O / rowsum -> bf16 -> global
// ViT's 257th token would cost a whole extra 128-row tile and a 17th K-step; instead the cls KEY is
// folded in on CUDA cores (one extra score per row, one rank-1 update of O) and the cls QUERY row is
// computed by a spare warp, so every tensor-core tile is exactly 128 x 256.
// hd = 88 is not a multiple of the UMMA K (16): Q's columns 88..95 are zeroed in shared memory, so
// whatever K holds there (the next head's finite values) contributes nothing.
#include <stdlib.h>
#include "common.cuh"
#include "ops.h"

namespace cgpt {

struct UmmaAttnParams {
  const __nv_bfloat16* q; long long ldq; int q_rows_per_batch;
  const __nv_bfloat16* k; const __nv_bfloat16* v; long long ldk, ldv; int kv_rows_per_batch;
  __nv_bfloat16* o; long long ldo;
  int H, head_dim, Tq, Tk, E, causal;  // Tq / Tk include the E extra (cls) row handled on CUDA cores
  float scale_log2e;
  int n_qtiles, nk_pad;                // 128-row query tiles; main keys padded to a multiple of 16
  int tmem_cols, o_col1;               // TMEM allocation; column of O for q-tile 0 when it does not alias S
  int q_bytes, k_region;               // smem carve-up
  int prefetch_distance;               // != 0: L2 prefetch for the item two rounds ahead of the one being loaded
  int n_items;                         // B * H (batch, head) items, walked by a persistent grid
  long long* dbg;                      // optional per-CTA phase timestamps [grid][16] (null in production)
};

constexpr int UA_THREADS = 384;
constexpr int SUB = 16384;  // one 128-row x 64-col bf16 sub-tile

__device__ __forceinline__ uint64_t make_desc_kmajor(uint32_t addr) { return make_smem_desc_sw128(addr); }
// MN-major operand, 128B swizzle: LBO = byte stride between 64-element MN atoms, SBO = 8-row K-group stride
__device__ __forceinline__ uint64_t make_desc_mnmajor(uint32_t addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>(1024u >> 4) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}
__device__ __forceinline__ float ua_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(UA_THREADS, 1)
attn_umma_kernel(const __grid_constant__ CUtensorMap map_q0, const __grid_constant__ CUtensorMap map_q1,
                 const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_k1,
                 const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_v1,
                 UmmaAttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by pointer arithmetic on the __shared__ array: a round trip through uintptr_t would
  // lose the address space and turn every shared-memory access below into a generic LD.E / ST.E
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;                       // n_qtiles x [sub0 | sub1]
  uint8_t* sK = smem + p.q_bytes;           // [sub0 | sub1], sub stride nk_pad*128; later P of q-tile 0
  uint8_t* sV = sK + p.k_region;            // [sub0 | sub1]
  const int kv_sub = p.nk_pad * 128;
  uint8_t* tail = sV + 2 * kv_sub;
  // Every barrier completes exactly once per item, so item `it` of this CTA waits with parity it & 1
  // (bar_xr: once per use of an x-buffer, parity (it >> 1) & 1).
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail);
  uint64_t* bar_qk = bars;       // TMA: Q + K landed
  uint64_t* bar_v = bars + 1;    // TMA: V landed
  uint64_t* bar_s = bars + 2;    // [2] S_qt in TMEM
  uint64_t* bar_p = bars + 4;    // [2] P_qt in smem (128 arrivals)
  uint64_t* bar_o = bars + 6;    // [2] O_qt in TMEM
  uint64_t* bar_x = bars + 8;    // extra-query scores written (one arrival per softmax thread)
  uint64_t* bar_tf = bars + 9;   // TMEM + Q pad columns released by the softmax threads (one arrival each)
  uint64_t* bar_vf = bars + 10;  // V tile released by the extra-query warps (64 arrivals)
  uint64_t* bar_xr = bars + 11;  // [2] extra key / value / query of an item staged in x-buffer [it & 1] (32 arrivals)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 13);
  // extra (cls) row staging, double-buffered by item parity: [xk 128 | xv 128 | xq 128 | xs 320] fp32
  constexpr int XBUF = 3 * 128 + 320;
  float* xbase = reinterpret_cast<float*>(tail + 128);
  float* xo = xbase + 2 * XBUF;                         // extra-query partial outputs [5][96]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#define UA_STAMP(slot) do { if (p.dbg) p.dbg[blockIdx.x * 16 + (slot)] = clock64(); } while (0)
  const int hd = p.head_dim, E = p.E;
  const int Tk_main = p.Tk - E, Tq_main = p.Tq - E;
  const int ksteps_s = (hd + 15) / 16;
  const int ksteps_o = p.nk_pad / 16;
  const int grid = static_cast<int>(gridDim.x);
  const int n_soft = 128 * p.n_qtiles;

  // all TMA loads of one (batch, head) item + an L2 prefetch for the item two rounds ahead
  auto issue_loads = [&](int item) {
    const int h = item % p.H, b = item / p.H;
    const int hcol = h * hd;
    const int qrow = b * p.q_rows_per_batch + E, krow = b * p.kv_rows_per_batch + E;
    const int q1_cols = hd - 64;
    const uint32_t q_tx = static_cast<uint32_t>(p.n_qtiles) * (64 * 128 * 2 + q1_cols * 128 * 2);
    // second sub-tile boxes: K needs columns up to roundup16(hd) (Q is zero beyond hd), V only hd
    const int k1_cols = ((hd + 15) & ~15) - 64, v1_cols = hd - 64;
    const uint32_t k_tx = static_cast<uint32_t>(64 + k1_cols) * p.nk_pad * 2;
    const uint32_t v_tx = static_cast<uint32_t>(64 + v1_cols) * p.nk_pad * 2;
    mbar_arrive_expect_tx(bar_qk, q_tx + k_tx);
    for (int qt = 0; qt < p.n_qtiles; ++qt) {
      tma_load_2d(sQ + qt * 2 * SUB, &map_q0, bar_qk, hcol, qrow + qt * 128);
      tma_load_2d(sQ + qt * 2 * SUB + SUB, &map_q1, bar_qk, hcol + 64, qrow + qt * 128);
    }
    tma_load_2d(sK, &map_k, bar_qk, hcol, krow);
    tma_load_2d(sK + kv_sub, &map_k1, bar_qk, hcol + 64, krow);
    mbar_arrive_expect_tx(bar_v, v_tx);
    tma_load_2d(sV, &map_v, bar_v, hcol, krow);
    tma_load_2d(sV + kv_sub, &map_v1, bar_v, hcol + 64, krow);
    const int ahead = item + 2 * grid;
    if (p.prefetch_distance > 0 && ahead < p.n_items) {
      const int hb = ahead / p.H, pc = (ahead % p.H) * hd;
      const int pq = hb * p.q_rows_per_batch + E, pk = hb * p.kv_rows_per_batch + E;
      for (int qt = 0; qt < p.n_qtiles; ++qt) {
        tma_prefetch_l2_2d(&map_q0, pc, pq + qt * 128);
        tma_prefetch_l2_2d(&map_q1, pc + 64, pq + qt * 128);
      }
      tma_prefetch_l2_2d(&map_k, pc, pk);
      tma_prefetch_l2_2d(&map_k1, pc + 64, pk);
      tma_prefetch_l2_2d(&map_v, pc, pk);
      tma_prefetch_l2_2d(&map_v1, pc + 64, pk);
    }
  };
  // extra (cls) key / value / query of one item's head as fp32 in an x-buffer (one 16-byte load per lane and row)
  auto stage_extra = [&](int item, float* xb) {
    const int h = item % p.H, b = item / p.H;
    const int hcol = h * hd;
    const __nv_bfloat16* kx = p.k + static_cast<long long>(b) * p.kv_rows_per_batch * p.ldk + hcol;
    const __nv_bfloat16* vx = p.v + static_cast<long long>(b) * p.kv_rows_per_batch * p.ldv + hcol;
    const __nv_bfloat16* qx = p.q + static_cast<long long>(b) * p.q_rows_per_batch * p.ldq + hcol;
    float* xk = xb; float* xv = xb + 128; float* xq = xb + 256;
    if (lane < 16) {
      const bool ok = lane * 8 < hd;
      uint4 a = make_uint4(0, 0, 0, 0), bq = a, c = a;
      if (ok) {
        a = *reinterpret_cast<const uint4*>(kx + lane * 8);
        bq = *reinterpret_cast<const uint4*>(vx + lane * 8);
        c = *reinterpret_cast<const uint4*>(qx + lane * 8);
      }
      const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {bq.x, bq.y, bq.z, bq.w}, cw[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        xk[lane * 8 + 2 * i] = bf16_lo(aw[i]); xk[lane * 8 + 2 * i + 1] = bf16_hi(aw[i]);
        xv[lane * 8 + 2 * i] = bf16_lo(bw[i]); xv[lane * 8 + 2 * i + 1] = bf16_hi(bw[i]);
        xq[lane * 8 + 2 * i] = bf16_lo(cw[i]) * p.scale_log2e; xq[lane * 8 + 2 * i + 1] = bf16_hi(cw[i]) * p.scale_log2e;
      }
    }
  };
  // zero Q's pad columns [hd, roundup16(hd)) of the second sub-tile: one 16-byte chunk per row
  // (disjoint from the bytes the TMA box writes; made visible to the tensor core by the proxy fence)
  auto zero_q_pad = [&](int qt, int r) {
    const int chunk = (hd - 64) >> 3;
    uint8_t* dst = sQ + qt * 2 * SUB + SUB + r * 128 + ((chunk ^ (r & 7)) << 4);
    *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
    fence_proxy_async();
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_q0); tma_prefetch_desc(&map_q1); tma_prefetch_desc(&map_k); tma_prefetch_desc(&map_v);
    mbar_init(bar_qk, 1);
    mbar_init(bar_v, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&bar_s[i], 1); mbar_init(&bar_p[i], 128); mbar_init(&bar_o[i], 1); }
    mbar_init(bar_x, n_soft);
    mbar_init(bar_tf, n_soft);
    mbar_init(bar_vf, 64);
    mbar_init(&bar_xr[0], 32);
    mbar_init(&bar_xr[1], 32);
    fence_barrier_init();
    fence_proxy_async();
    issue_loads(static_cast<int>(blockIdx.x));
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, p.tmem_cols);
    tmem_relinquish();
  }
  if (warp >= 4 && (hd & 15)) {
    const int qt = (warp - 4) >> 2;
    if (qt < p.n_qtiles) zero_q_pad(qt, (warp & 3) * 32 + lane);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t o_col0 = p.n_qtiles == 2 ? 0u : static_cast<uint32_t>(p.o_col1);

  int it = 0;
  for (int item = static_cast<int>(blockIdx.x); item < p.n_items; item += grid, ++it) {
    const uint32_t ph = it & 1;
    const int h = item % p.H, b = item / p.H;
    const int hcol = h * hd;
    const long long qrow_base = static_cast<long long>(b) * p.q_rows_per_batch;
    float* xb = xbase + (it & 1) * XBUF;
    float* xk = xb; float* xv = xb + 128; float* xq = xb + 256; float* xs = xb + 384;
    if (threadIdx.x == 0) UA_STAMP(0);      // debug stamps are per item (the last item of the CTA survives)
    if (threadIdx.x == 128) UA_STAMP(15);

    if (warp == 0) {
      if (lane == 0) {
        // ---------------------------------------------------------------- MMA issue
        const uint32_t idesc_s = make_idesc_bf16(128, p.nk_pad);
        const uint32_t idesc_o = make_idesc_bf16(128, 128) | (1u << 16);   // B (= V) is MN-major
        mbar_wait(bar_qk, ph);
        if (it > 0) mbar_wait(bar_tf, ph ^ 1);   // previous item's O drained from TMEM, Q pad columns re-zeroed
        UA_STAMP(2);
        tcgen05_fence_after();
        for (int qt = 0; qt < p.n_qtiles; ++qt) {
          const uint32_t d_tmem = tmem_base + qt * 256;
          for (int ks = 0; ks < ksteps_s; ++ks) {
            const uint64_t ad = make_desc_kmajor(smem_u32(sQ + qt * 2 * SUB + (ks >> 2) * SUB)) + 2 * (ks & 3);
            const uint64_t bd = make_desc_kmajor(smem_u32(sK + (ks >> 2) * kv_sub)) + 2 * (ks & 3);
            umma_bf16(d_tmem, ad, bd, idesc_s, ks != 0);
          }
          umma_commit(&bar_s[qt]);
        }
        UA_STAMP(3);
        mbar_wait(bar_v, ph);
        for (int qt = 0; qt < p.n_qtiles; ++qt) {
          mbar_wait(&bar_p[qt], ph);
          UA_STAMP(4 + qt);
          tcgen05_fence_after();
          const uint8_t* sP = qt == 0 ? sK : sQ;
          const uint32_t d_tmem = tmem_base + (qt == 0 ? o_col0 : 256u);
          for (int ks = 0; ks < ksteps_o; ++ks) {
            const uint64_t ad = make_desc_kmajor(smem_u32(sP + (ks >> 2) * SUB)) + 2 * (ks & 3);
            const uint64_t bd = make_desc_mnmajor(smem_u32(sV + ks * 16 * 128), static_cast<uint32_t>(kv_sub));
            umma_bf16(d_tmem, ad, bd, idesc_o, ks != 0);
          }
          umma_commit(&bar_o[qt]);
        }
        // next item's loads as soon as the tensor core (P / V) and the extra-query warps (V) are done with
        // this item's tiles: they overlap the epilogue of this item
        if (item + grid < p.n_items) {
          mbar_wait(&bar_o[p.n_qtiles - 1], ph);
          if (E) mbar_wait(bar_vf, ph);
          issue_loads(item + grid);
        }
        UA_STAMP(13);
      }
    } else if (warp == 2 || warp == 3) {
      // ---------------------------------------------------------------- extra (cls) query row, CUDA cores
      // scores against the main keys come from the softmax threads (thread j <-> key j, xs[1 + j]);
      // softmax over the Tk scores by warp 2, then P.V as a GEMV out of the V tile in shared memory
      if (E) {
        if (warp == 3) {
          // stage this item's extra key / value / query; the buffer was last used two items ago
          stage_extra(item, xb);
          mbar_arrive(&bar_xr[it & 1]);
        }
        const int t64 = threadIdx.x - 64;   // 0..63
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
          for (int j = lane; j < p.Tk; j += 32) mx = fmaxf(mx, xs[j]);
          mx = warp_max(mx);
          float sum = 0.f;
          for (int j = lane; j < p.Tk; j += 32) {
            const float e = ua_exp2(xs[j] - mx);
            xs[j] = e;
            sum += e;
          }
          sum = warp_sum(sum);
          if (lane == 0) xs[p.Tk] = 1.f / sum;
        }
        mbar_wait(bar_v, ph);
        asm volatile("bar.sync 2, 64;" ::: "memory");
        // 11 (hd/8) column chunks x 5 key partitions = 55 threads
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
          const float inv = xs[p.Tk];
          float f[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int d = t64 * 8 + i;
            f[i] = (xo[d] + xo[96 + d] + xo[192 + d] + xo[288 + d] + xo[384 + d] + xs[0] * xv[d]) * inv;
          }
          *reinterpret_cast<uint4*>(p.o + qrow_base * p.ldo + hcol + t64 * 8) =
              make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
        }
        asm volatile("bar.sync 2, 64;" ::: "memory");   // xo is reused by the next item
      }
    } else if (warp >= 4) {
      // ---------------------------------------------------------------- softmax + epilogue, thread = row
      const int qt = (warp - 4) >> 2;
      if (qt < p.n_qtiles) {
        const int r = (warp & 3) * 32 + lane;
        const int q_main = qt * 128 + r;                 // index among the main query rows
        const bool row_ok = q_main < Tq_main;
        const int q_abs = E + q_main;
        const int offs = p.Tk - p.Tq;
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
        // scores against the extra key (this row) and of the extra query against main key q_main
        // (thread <-> key), read from the TMA-loaded, swizzled Q / K tiles in shared memory while the
        // S products run; both tiles are still intact (P is written only after pass 1)
        float s_x = -INFINITY;
        if (E) {
          mbar_wait(&bar_xr[it & 1], (it >> 1) & 1);
          mbar_wait(bar_qk, ph);
          float a1 = 0.f, a2 = 0.f;
          const int nch = hd >> 3;
          const uint8_t* qrow = sQ + qt * 2 * SUB + r * 128;
          const uint8_t* krow = sK + q_main * 128;        // key index == q_main (requires q_main < nk_pad)
          const bool kok = q_main < Tk_main;
          for (int c = 0; c < nch; ++c) {
            const int sub = c >> 3, cc = c & 7;
            const uint4 qv = *reinterpret_cast<const uint4*>(qrow + sub * SUB + ((cc ^ (r & 7)) << 4));
            const float4 k0 = *reinterpret_cast<const float4*>(xk + c * 8), k1 = *reinterpret_cast<const float4*>(xk + c * 8 + 4);
            a1 += bf16_lo(qv.x) * k0.x + bf16_hi(qv.x) * k0.y + bf16_lo(qv.y) * k0.z + bf16_hi(qv.y) * k0.w +
                  bf16_lo(qv.z) * k1.x + bf16_hi(qv.z) * k1.y + bf16_lo(qv.w) * k1.z + bf16_hi(qv.w) * k1.w;
            if (kok) {
              const uint4 kv = *reinterpret_cast<const uint4*>(krow + sub * kv_sub + ((cc ^ (q_main & 7)) << 4));
              const float4 q0 = *reinterpret_cast<const float4*>(xq + c * 8), q1 = *reinterpret_cast<const float4*>(xq + c * 8 + 4);
              a2 += bf16_lo(kv.x) * q0.x + bf16_hi(kv.x) * q0.y + bf16_lo(kv.y) * q0.z + bf16_hi(kv.y) * q0.w +
                    bf16_lo(kv.z) * q1.x + bf16_hi(kv.z) * q1.y + bf16_lo(kv.w) * q1.z + bf16_hi(kv.w) * q1.w;
            }
          }
          if (row_ok) s_x = a1;
          if (kok) xs[1 + q_main] = a2;
          mbar_arrive(bar_x);
        }
        if (threadIdx.x == 128) UA_STAMP(6);
        mbar_wait(&bar_s[qt], ph);
        if (threadIdx.x == 128) UA_STAMP(7);
        tcgen05_fence_after();
        const uint32_t s_addr = t_lane + qt * 256;
        // pass 1: row max of the raw scores (mask: padded keys, causal).  16-column TMEM loads, the next one
        // in flight while the current is reduced; full chunks skip the mask.
        const int kmax = p.causal ? min(Tk_main, q_abs + offs + 1 - E) : Tk_main;   // main keys [0, kmax) visible
        float mx = s_x;
        {
          uint32_t va[16], vb[16];
          tmem_ld_x16(s_addr, va);
#pragma unroll 1
          for (int c = 0; c < p.nk_pad; c += 32) {
            tmem_ld_wait();
            if (c + 16 < p.nk_pad) tmem_ld_x16(s_addr + c + 16, vb);
            if (c + 16 <= kmax) {
#pragma unroll
              for (int i = 0; i < 16; ++i) mx = fmaxf(mx, __uint_as_float(va[i]));
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) if (c + i < kmax) mx = fmaxf(mx, __uint_as_float(va[i]));
            }
            if (c + 16 < p.nk_pad) {
              tmem_ld_wait();
              if (c + 32 < p.nk_pad) tmem_ld_x16(s_addr + c + 32, va);
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
        if (threadIdx.x == 128) UA_STAMP(8);
        const float neg_ms = -mx * p.scale_log2e;
        // P may only overwrite K / Q once BOTH S products have been issued and retired
        for (int t = 0; t < p.n_qtiles; ++t) mbar_wait(&bar_s[t], ph);
        uint8_t* sP = qt == 0 ? sK : sQ;
        float sum = 0.f;
        auto emit = [&](const uint32_t* v, int c) {
          float e[16];
          if (c + 16 <= kmax) {
#pragma unroll
            for (int i = 0; i < 16; ++i) e[i] = ua_exp2(fmaf(__uint_as_float(v[i]), p.scale_log2e, neg_ms));
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              e[i] = (c + i < kmax) ? ua_exp2(fmaf(__uint_as_float(v[i]), p.scale_log2e, neg_ms)) : 0.f;
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) sum += e[i];
          // two 16-byte chunks of the K-major, 128B-swizzled P tile
          const int st = c >> 6, j0 = (c & 63) >> 3;
          uint8_t* rowp = sP + st * SUB + r * 128;
          *reinterpret_cast<uint4*>(rowp + (((j0) ^ (r & 7)) << 4)) =
              make_uint4(pack_bf16x2(e[0], e[1]), pack_bf16x2(e[2], e[3]), pack_bf16x2(e[4], e[5]), pack_bf16x2(e[6], e[7]));
          *reinterpret_cast<uint4*>(rowp + (((j0 + 1) ^ (r & 7)) << 4)) =
              make_uint4(pack_bf16x2(e[8], e[9]), pack_bf16x2(e[10], e[11]), pack_bf16x2(e[12], e[13]), pack_bf16x2(e[14], e[15]));
        };
        {
          uint32_t va[16], vb[16];
          tmem_ld_x16(s_addr, va);
#pragma unroll 1
          for (int c = 0; c < p.nk_pad; c += 32) {
            tmem_ld_wait();
            if (c + 16 < p.nk_pad) tmem_ld_x16(s_addr + c + 16, vb);
            emit(va, c);
            if (c + 16 < p.nk_pad) {
              tmem_ld_wait();
              if (c + 32 < p.nk_pad) tmem_ld_x16(s_addr + c + 32, va);
              emit(vb, c + 16);
            }
          }
        }
        float p_x = 0.f;
        if (E && row_ok) {
          p_x = ua_exp2(fmaf(s_x, p.scale_log2e, neg_ms));
          sum += p_x;
        }
        fence_proxy_async();        // P (generic-proxy stores) -> visible to the tensor core
        tcgen05_fence_before();     // this thread's TMEM reads of S are done before O may overwrite them
        mbar_arrive(&bar_p[qt]);
        if (threadIdx.x == 128) UA_STAMP(9);
        // ---- epilogue
        mbar_wait(&bar_o[qt], ph);
        if (threadIdx.x == 128) UA_STAMP(10);
        tcgen05_fence_after();
        const float inv = sum > 0.f ? 1.f / sum : 0.f;
        const uint32_t o_addr = t_lane + (qt == 0 ? o_col0 : 256u);
        __nv_bfloat16* orow = p.o + (qrow_base + q_abs) * p.ldo + hcol;
        auto store16 = [&](const uint32_t* v, int c) {
          if (!row_ok) return;
          float f[16];
          if (E) {
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
              const float4 x4 = *reinterpret_cast<const float4*>(xv + ((c + i) & 127));
              f[i] = (__uint_as_float(v[i]) + p_x * x4.x) * inv;
              f[i + 1] = (__uint_as_float(v[i + 1]) + p_x * x4.y) * inv;
              f[i + 2] = (__uint_as_float(v[i + 2]) + p_x * x4.z) * inv;
              f[i + 3] = (__uint_as_float(v[i + 3]) + p_x * x4.w) * inv;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]) * inv;
          }
          *reinterpret_cast<uint4*>(orow + c) =
              make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
          if (c + 8 < hd)
            *reinterpret_cast<uint4*>(orow + c + 8) =
                make_uint4(pack_bf16x2(f[8], f[9]), pack_bf16x2(f[10], f[11]), pack_bf16x2(f[12], f[13]), pack_bf16x2(f[14], f[15]));
        };
        {
          uint32_t va[16], vb[16];
          tmem_ld_x16(o_addr, va);
#pragma unroll 1
          for (int c = 0; c < hd; c += 32) {
            tmem_ld_wait();
            if (c + 16 < hd) tmem_ld_x16(o_addr + c + 16, vb);
            store16(va, c);
            if (c + 16 < hd) {
              tmem_ld_wait();
              if (c + 32 < hd) tmem_ld_x16(o_addr + c + 32, va);
              store16(vb, c + 16);
            }
          }
        }
        if (threadIdx.x == 128) UA_STAMP(11);
        if (item + grid < p.n_items) {
          // the next item's S products need Q's pad columns zero again: P of q-tile 1 was written over the Q
          // tiles, so wait until the last P.V product has consumed it, then restore this row's chunk
          if ((hd & 15) && p.n_qtiles == 2) {
            mbar_wait(&bar_o[1], ph);
            zero_q_pad(qt, r);
          }
          tcgen05_fence_before();   // this thread's TMEM reads of O are done before the next S may overwrite them
          mbar_arrive(bar_tf);
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
#undef UA_STAMP
}

// ------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn2)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn2 encode_fn() {
  static EncodeTiledFn2 fn = nullptr;
  if (fn) return fn;
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn2>(ptr);
  return fn;
}
static int make_map(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld, int box_cols,
                    int box_rows) {
  EncodeTiledFn2 fn = encode_fn();
  CGPT_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled driver entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CGPT_REQUIRE(r == CUDA_SUCCESS, "attention: cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld box=%dx%d",
               (int)r, rows, cols, ld, box_cols, box_rows);
  return 0;
}

// 1 = this shape is served by the tcgen05 kernel
int attn_umma_supported(const cgpt_attn_args* a) {
  if (a->decode_kernel != 0 || a->P != 0) return 0;
  if (a->head_dim <= 64 || a->head_dim > 128 || (a->head_dim & 7)) return 0;
  if (a->B * (long long)a->H > 0x7fffffffLL) return 0;
  int E = 0;
  if (a->Tk > 256) {
    if (a->Tk - 1 > 256 || a->Tq != a->Tk || a->causal) return 0;
    E = 1;
  }
  if (a->Tq - E > 256 || a->Tq - E < 1) return 0;
  if ((reinterpret_cast<uintptr_t>(a->q) | reinterpret_cast<uintptr_t>(a->k) | reinterpret_cast<uintptr_t>(a->v)) & 15) return 0;
  return 1;
}

int attention_umma(const cgpt_attn_args* a, cudaStream_t stream) {
  UmmaAttnParams p;
  p.q = (const __nv_bfloat16*)a->q; p.ldq = a->ldq; p.q_rows_per_batch = a->q_rows_per_batch;
  p.k = (const __nv_bfloat16*)a->k; p.v = (const __nv_bfloat16*)a->v; p.ldk = a->ldk; p.ldv = a->ldv;
  p.kv_rows_per_batch = a->kv_rows_per_batch;
  p.o = (__nv_bfloat16*)a->o; p.ldo = a->ldo;
  p.H = a->H; p.head_dim = a->head_dim; p.Tq = a->Tq; p.Tk = a->Tk; p.causal = a->causal;
  p.E = a->Tk > 256 ? 1 : 0;
  p.scale_log2e = a->scale * 1.4426950408889634f;
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  p.prefetch_distance = getenv("CGPT_ATTN_NO_PREFETCH") ? 0 : 1;
  p.dbg = reinterpret_cast<long long*>(getenv("CGPT_ATTN_DBG") ? strtoull(getenv("CGPT_ATTN_DBG"), nullptr, 0) : 0ull);
  const int tq_main = a->Tq - p.E, tk_main = a->Tk - p.E;
  p.n_qtiles = (tq_main + 127) / 128;
  p.nk_pad = (tk_main + 15) / 16 * 16;
  if (p.n_qtiles == 2) { p.tmem_cols = 512; p.o_col1 = 0; }
  else if (p.nk_pad <= 128) { p.tmem_cols = 256; p.o_col1 = 128; }
  else { p.tmem_cols = 512; p.o_col1 = 256; }
  p.q_bytes = p.n_qtiles * 2 * SUB;
  const int kv_sub = p.nk_pad * 128;
  const int p_bytes = (p.nk_pad + 63) / 64 * SUB;
  p.k_region = 2 * kv_sub > p_bytes ? 2 * kv_sub : p_bytes;
  const int smem = p.q_bytes + p.k_region + 2 * kv_sub + 128 + 2 * (3 * 128 + 320) * 4 + 5 * 96 * 4 + 1024;
  CGPT_REQUIRE(smem <= 227 * 1024, "attention_umma: shared memory %d too large", smem);

  const long long q_rows = (long long)a->B * a->q_rows_per_batch, kv_rows = (long long)a->B * a->kv_rows_per_batch;
  const long long cols = (long long)a->H * a->head_dim;
  CUtensorMap mq0, mq1, mk, mk1, mv, mv1;
  if (int rc = make_map(&mq0, a->q, q_rows, cols, a->ldq, 64, 128)) return rc;
  if (int rc = make_map(&mq1, a->q, q_rows, cols, a->ldq, a->head_dim - 64, 128)) return rc;
  if (int rc = make_map(&mk, a->k, kv_rows, cols, a->ldk, 64, p.nk_pad)) return rc;
  if (int rc = make_map(&mv, a->v, kv_rows, cols, a->ldv, 64, p.nk_pad)) return rc;
  if (int rc = make_map(&mk1, a->k, kv_rows, cols, a->ldk, ((a->head_dim + 15) & ~15) - 64, p.nk_pad)) return rc;
  if (int rc = make_map(&mv1, a->v, kv_rows, cols, a->ldv, a->head_dim - 64, p.nk_pad)) return rc;
  static int configured_smem = 0;
  if (smem > configured_smem) {
    CGPT_CHECK_CUDA(cudaFuncSetAttribute(attn_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured_smem = smem;
  }
  // persistent grid: one CTA per SM walks items blockIdx.x, blockIdx.x + grid, ...
  p.n_items = a->B * a->H;
  static const bool one_shot = getenv("CGPT_ATTN_ONE_SHOT") != nullptr;   // A/B: one CTA per item, as before
  int grid = sms;   // 384 threads x 154 registers: one resident CTA per SM
  if (one_shot || grid > p.n_items) grid = p.n_items;
  attn_umma_kernel<<<grid, UA_THREADS, smem, stream>>>(mq0, mq1, mk, mk1, mv, mv1, p);
  CGPT_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace cgpt
