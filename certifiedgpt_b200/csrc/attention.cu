// Attention cores of the three towers on the smoothing path:
//   - EVA ViT self-attention, T=257, 16 heads x 88, no mask        (eva_vit.py:133-150)
//   - Q-Former self (32x32) and cross (32x257) attention, 12 x 64    (Qformer.py:195-275)
//   - Llama causal prefill with a batch-shared prompt-prefix K/V     (HF LlamaAttention)
//   - Llama single-token decode against the per-sample KV cache
//
// flash_attn_kernel: one CTA = 64 query rows of one (batch, head); 4 warps x 16 rows;
// K/V streamed in 64-row tiles through a double-buffered cp.async pipeline; QK^T and PV on
// mma.sync.m16n8k16 bf16 with fp32 online softmax.  Attention is 2.8 % of the ViT FLOPs at
// 224 px, so this legacy-tensor-path kernel is not on the roofline-critical path; the GEMMs are.
// The K/V row index space is [shared prefix rows | per-batch rows]: rows < P are read from a
// batch-invariant buffer (the "<s>[INST] <Img>" prompt prefix), the rest from this batch's rows.
#include <stdlib.h>
#include "common.cuh"
#include "ops.h"

namespace cgpt {

struct AttnParams {
  const __nv_bfloat16* q; long long ldq; int q_rows_per_batch;
  const __nv_bfloat16* k; const __nv_bfloat16* v; long long ldk, ldv; int kv_rows_per_batch;
  const __nv_bfloat16* kp; const __nv_bfloat16* vp; long long ldkp, ldvp; int P;
  __nv_bfloat16* o; long long ldo;
  int Tq, Tk;          // valid query rows / total key rows (prefix included)
  int head_dim;        // real head dim (64, 88, 128)
  float scale_log2e;   // softmax scale * log2(e)
  int causal;          // query i sees keys <= i + (Tk - Tq)
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

constexpr int FA_BN = 64;   // key/value tile rows; the query tile is 16 rows per warp (NW warps)

template <int HD, int ROWS, int THREADS>
__device__ __forceinline__ void load_rows_tile(uint32_t smem_tile, const __nv_bfloat16* base, long long ld,
                                               int row0, int nrows_valid, int head_dim,
                                               const __nv_bfloat16* pbase, long long pld, int P) {
  // ROWS rows x HD cols, 16-byte chunks; rows >= nrows_valid and cols >= head_dim are zero-filled
  constexpr int CH = HD / 8;
  constexpr int LDS = HD + 8;
  for (int i = threadIdx.x; i < ROWS * CH; i += THREADS) {
    const int r = i / CH, c = i - r * CH;
    const int row = row0 + r;
    const bool valid = row < nrows_valid && c * 8 < head_dim;
    const __nv_bfloat16* src = base;
    if (valid) src = (row < P) ? pbase + row * pld + c * 8 : base + static_cast<long long>(row - P) * ld + c * 8;
    cp_async16(smem_tile + (r * LDS + c * 8) * 2, src, valid);
  }
}

template <int HD, int NW>
__global__ void __launch_bounds__(NW * 32) flash_attn_kernel(AttnParams p) {
  constexpr int FA_BM = 16 * NW;
  constexpr int FA_THREADS = 32 * NW;
  constexpr int LDS = HD + 8;
  constexpr int KSTEPS = HD / 16;
  constexpr int DTILES = HD / 8;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const uint32_t sQ = smem_u32(smem_raw);
  const uint32_t sK = sQ + FA_BM * LDS * 2;           // 2 buffers
  const uint32_t sV = sK + 2 * FA_BN * LDS * 2;       // 2 buffers
  const int qb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;

  const __nv_bfloat16* qbase = p.q + static_cast<long long>(b) * p.q_rows_per_batch * p.ldq + h * p.head_dim;
  const __nv_bfloat16* kbase = p.k + static_cast<long long>(b) * p.kv_rows_per_batch * p.ldk + h * p.head_dim;
  const __nv_bfloat16* vbase = p.v + static_cast<long long>(b) * p.kv_rows_per_batch * p.ldv + h * p.head_dim;
  const __nv_bfloat16* kpbase = p.kp ? p.kp + h * p.head_dim : nullptr;
  const __nv_bfloat16* vpbase = p.vp ? p.vp + h * p.head_dim : nullptr;

  const int q0 = qb * FA_BM;
  const int offs = p.Tk - p.Tq;
  int n_tiles = (p.Tk + FA_BN - 1) / FA_BN;
  if (p.causal) {
    const int last_q = min(q0 + FA_BM, p.Tq) - 1;
    n_tiles = min(n_tiles, (last_q + offs) / FA_BN + 1);
  }

  load_rows_tile<HD, FA_BM, FA_THREADS>(sQ, qbase, p.ldq, q0, p.Tq, p.head_dim, nullptr, 0, 0);
  load_rows_tile<HD, FA_BN, FA_THREADS>(sK, kbase, p.ldk, 0, p.Tk, p.head_dim, kpbase, p.ldkp, p.P);
  load_rows_tile<HD, FA_BN, FA_THREADS>(sV, vbase, p.ldv, 0, p.Tk, p.head_dim, vpbase, p.ldvp, p.P);
  cp_async_commit();

  uint32_t qf[KSTEPS][4];
  float o_acc[DTILES][4];
#pragma unroll
  for (int i = 0; i < DTILES; ++i) { o_acc[i][0] = o_acc[i][1] = o_acc[i][2] = o_acc[i][3] = 0.f; }
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  const bool warp_active = q0 + warp * 16 < p.Tq;

  for (int j = 0; j < n_tiles; ++j) {
    const int buf = j & 1;
    if (j + 1 < n_tiles) {
      load_rows_tile<HD, FA_BN, FA_THREADS>(sK + (buf ^ 1) * FA_BN * LDS * 2, kbase, p.ldk, (j + 1) * FA_BN, p.Tk, p.head_dim,
                         kpbase, p.ldkp, p.P);
      load_rows_tile<HD, FA_BN, FA_THREADS>(sV + (buf ^ 1) * FA_BN * LDS * 2, vbase, p.ldv, (j + 1) * FA_BN, p.Tk, p.head_dim,
                         vpbase, p.ldvp, p.P);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (j == 0) {
#pragma unroll
      for (int ks = 0; ks < KSTEPS; ++ks) {
        const uint32_t addr = sQ + ((warp * 16 + (lane & 15)) * LDS + ks * 16 + (lane >> 4) * 8) * 2;
        ldmatrix_x4(addr, qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
      }
    }
    if (warp_active) {
      const uint32_t kt = sK + buf * FA_BN * LDS * 2;
      const uint32_t vt = sV + buf * FA_BN * LDS * 2;
      float s[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
      // keys valid in this tile (T=257 leaves ONE key in the 5th tile: skip its other MMAs)
      const int nvalid = min(FA_BN, p.Tk - j * FA_BN);
#pragma unroll
      for (int ks = 0; ks < KSTEPS; ++ks) {
#pragma unroll
        for (int np = 0; np < 4; ++np) {  // pairs of 8-wide key tiles
          if (np * 16 >= nvalid) continue;
          uint32_t b0, b1, b2, b3;
          const uint32_t addr = kt + ((np * 16 + (lane & 7) + (lane >> 4) * 8) * LDS + ks * 16 + ((lane >> 3) & 1) * 8) * 2;
          ldmatrix_x4(addr, b0, b1, b2, b3);
          mma_bf16_16816(s[2 * np], qf[ks], b0, b1);
          mma_bf16_16816(s[2 * np + 1], qf[ks], b2, b3);
        }
      }
      // online softmax on RAW scores (rows g and g+8 of this warp's 16); the softmax scale is folded
      // into the exp2 argument (one FFMA per element).  Masking code only runs on tiles that need it.
      const int qrow0 = q0 + warp * 16 + g;
      const bool need_mask = (j * FA_BN + FA_BN > p.Tk) ||
                             (p.causal && (j * FA_BN + FA_BN - 1 > q0 + warp * 16 + offs));
      if (need_mask) {
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int col = j * FA_BN + nt * 8 + 2 * t + (e & 1);
            const int qrow = qrow0 + (e >> 1) * 8;
            const bool ok = col < p.Tk && (!p.causal || col <= qrow + offs);
            if (!ok) s[nt][e] = -INFINITY;
          }
        }
      }
      float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        mx[0] = fmaxf(mx[0], fmaxf(s[nt][0], s[nt][1]));
        mx[1] = fmaxf(mx[1], fmaxf(s[nt][2], s[nt][3]));
      }
      float corr[2], neg_ms[2];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
        mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
        const float m_new = fmaxf(m_run[r], mx[r]);
        const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
        corr[r] = fast_exp2((m_run[r] - m_use) * p.scale_log2e);
        neg_ms[r] = -m_use * p.scale_log2e;
        m_run[r] = m_new;
        l_run[r] *= corr[r];
      }
      float rs[2] = {0.f, 0.f};
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float pv = fast_exp2(fmaf(s[nt][e], p.scale_log2e, neg_ms[e >> 1]));
          s[nt][e] = pv;
          rs[e >> 1] += pv;
        }
      }
      l_run[0] += rs[0];
      l_run[1] += rs[1];
      if (corr[0] != 1.f || corr[1] != 1.f) {   // (per-thread; skipping is exact when the max did not move)
#pragma unroll
        for (int dt = 0; dt < DTILES; ++dt) {
          o_acc[dt][0] *= corr[0]; o_acc[dt][1] *= corr[0];
          o_acc[dt][2] *= corr[1]; o_acc[dt][3] *= corr[1];
        }
      }
      // O += P V
#pragma unroll
      for (int kk = 0; kk < FA_BN / 16; ++kk) {
        if (kk * 16 >= nvalid) continue;
        uint32_t pa[4];
        pa[0] = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]);
        pa[1] = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
        pa[2] = pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
        pa[3] = pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
        for (int dp = 0; dp < DTILES / 2; ++dp) {
          uint32_t b0, b1, b2, b3;
          const uint32_t addr = vt + ((kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDS + dp * 16 + (lane >> 4) * 8) * 2;
          ldmatrix_x4_trans(addr, b0, b1, b2, b3);
          mma_bf16_16816(o_acc[2 * dp], pa, b0, b1);
          mma_bf16_16816(o_acc[2 * dp + 1], pa, b2, b3);
        }
      }
    }
    __syncthreads();
  }

  if (warp_active) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
      l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
    }
    const float inv0 = l_run[0] > 0.f ? 1.f / l_run[0] : 0.f;
    const float inv1 = l_run[1] > 0.f ? 1.f / l_run[1] : 0.f;
    const int qrow0 = q0 + warp * 16 + g;
    __nv_bfloat16* obase = p.o + static_cast<long long>(b) * p.q_rows_per_batch * p.ldo + h * p.head_dim;
#pragma unroll
    for (int dt = 0; dt < DTILES; ++dt) {
      const int col = dt * 8 + 2 * t;
      if (col < p.head_dim) {
        if (qrow0 < p.Tq)
          *reinterpret_cast<uint32_t*>(obase + static_cast<long long>(qrow0) * p.ldo + col) =
              pack_bf16x2(o_acc[dt][0] * inv0, o_acc[dt][1] * inv0);
        if (qrow0 + 8 < p.Tq)
          *reinterpret_cast<uint32_t*>(obase + static_cast<long long>(qrow0 + 8) * p.ldo + col) =
              pack_bf16x2(o_acc[dt][2] * inv1, o_acc[dt][3] * inv1);
      }
    }
  }
}

// ---------------------------------------------------------------- single-token decode
// One warp per (batch, head).  8 lanes cooperate on one K/V row (a 256-byte row of a 128-wide head is
// read as 8 x 32 B, fully coalesced), 4 rows per warp iteration; scores go through shared memory for the
// softmax, then the same mapping accumulates P.V.  HBM-bound: KV bytes per (batch, head) = 2*Tk*HD*2.
template <int N>
__device__ __forceinline__ void load_bf16_vec(const __nv_bfloat16* src, float* v) {
  if constexpr (N == 16) {
    const uint4 a = *reinterpret_cast<const uint4*>(src);
    const uint4 b = *reinterpret_cast<const uint4*>(src + 8);
    v[0] = bf16_lo(a.x); v[1] = bf16_hi(a.x); v[2] = bf16_lo(a.y); v[3] = bf16_hi(a.y);
    v[4] = bf16_lo(a.z); v[5] = bf16_hi(a.z); v[6] = bf16_lo(a.w); v[7] = bf16_hi(a.w);
    v[8] = bf16_lo(b.x); v[9] = bf16_hi(b.x); v[10] = bf16_lo(b.y); v[11] = bf16_hi(b.y);
    v[12] = bf16_lo(b.z); v[13] = bf16_hi(b.z); v[14] = bf16_lo(b.w); v[15] = bf16_hi(b.w);
  } else if constexpr (N == 8) {
    const uint4 a = *reinterpret_cast<const uint4*>(src);
    v[0] = bf16_lo(a.x); v[1] = bf16_hi(a.x); v[2] = bf16_lo(a.y); v[3] = bf16_hi(a.y);
    v[4] = bf16_lo(a.z); v[5] = bf16_hi(a.z); v[6] = bf16_lo(a.w); v[7] = bf16_hi(a.w);
  } else {
    const uint2 a = *reinterpret_cast<const uint2*>(src);
    v[0] = bf16_lo(a.x); v[1] = bf16_hi(a.x); v[2] = bf16_lo(a.y); v[3] = bf16_hi(a.y);
  }
}

template <int HD>
__global__ void __launch_bounds__(128) decode_attn_kernel(AttnParams p) {
  constexpr int MAX_CTX = 1024;
  constexpr int DPL = HD / 8;  // dims per lane: 8 lanes span one head row
  __shared__ float s_p[4][MAX_CTX];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane & 7, grp = lane >> 3;
  const int h = blockIdx.x * 4 + warp, b = blockIdx.y;
  const __nv_bfloat16* q = p.q + static_cast<long long>(b) * p.q_rows_per_batch * p.ldq + h * HD + sub * DPL;
  const __nv_bfloat16* kbase = p.k + static_cast<long long>(b) * p.kv_rows_per_batch * p.ldk + h * HD + sub * DPL;
  const __nv_bfloat16* vbase = p.v + static_cast<long long>(b) * p.kv_rows_per_batch * p.ldv + h * HD + sub * DPL;
  const __nv_bfloat16* kpb = p.kp ? p.kp + h * HD + sub * DPL : nullptr;
  const __nv_bfloat16* vpb = p.vp ? p.vp + h * HD + sub * DPL : nullptr;
  float qv[DPL];
  load_bf16_vec<DPL>(q, qv);
#pragma unroll
  for (int i = 0; i < DPL; ++i) qv[i] *= p.scale_log2e;
  float mx = -INFINITY;
  for (int pos0 = 0; pos0 < p.Tk; pos0 += 4) {
    const int pos = pos0 + grp;
    float acc = 0.f;
    if (pos < p.Tk) {
      const __nv_bfloat16* kr = pos < p.P ? kpb + static_cast<long long>(pos) * p.ldkp
                                          : kbase + static_cast<long long>(pos - p.P) * p.ldk;
      float kv[DPL];
      load_bf16_vec<DPL>(kr, kv);
#pragma unroll
      for (int i = 0; i < DPL; ++i) acc = fmaf(kv[i], qv[i], acc);
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    if (pos < p.Tk) {
      if (sub == 0) s_p[warp][pos] = acc;
      mx = fmaxf(mx, acc);
    }
  }
  mx = warp_max(mx);
  __syncwarp();
  float sum = 0.f;
  for (int pos = lane; pos < p.Tk; pos += 32) {
    const float e = fast_exp2(s_p[warp][pos] - mx);
    s_p[warp][pos] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  __syncwarp();
  float acc[DPL];
#pragma unroll
  for (int i = 0; i < DPL; ++i) acc[i] = 0.f;
  for (int pos = grp; pos < p.Tk; pos += 4) {
    const __nv_bfloat16* vr = pos < p.P ? vpb + static_cast<long long>(pos) * p.ldvp
                                        : vbase + static_cast<long long>(pos - p.P) * p.ldv;
    const float pr = s_p[warp][pos];
    float vv[DPL];
    load_bf16_vec<DPL>(vr, vv);
#pragma unroll
    for (int i = 0; i < DPL; ++i) acc[i] = fmaf(pr, vv[i], acc[i]);
  }
  const float inv = 1.f / sum;
#pragma unroll
  for (int i = 0; i < DPL; ++i) {
    acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 8);
    acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 16);
    acc[i] *= inv;
  }
  if (grp == 0) {
    __nv_bfloat16* o = p.o + static_cast<long long>(b) * p.q_rows_per_batch * p.ldo + h * HD + sub * DPL;
#pragma unroll
    for (int i = 0; i < DPL; i += 2) *reinterpret_cast<uint32_t*>(o + i) = pack_bf16x2(acc[i], acc[i + 1]);
  }
}

// ---------------------------------------------------------------- host
template <int HD, int NW>
static int launch_flash_nw(const AttnParams& p, int B, int H, cudaStream_t stream) {
  constexpr int BMQ = 16 * NW;
  constexpr int smem = (BMQ + 4 * FA_BN) * (HD + 8) * 2;
  static bool configured = false;
  if (!configured) {
    CGPT_CHECK_CUDA(cudaFuncSetAttribute(flash_attn_kernel<HD, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  dim3 grid((p.Tq + BMQ - 1) / BMQ, H, B);
  flash_attn_kernel<HD, NW><<<grid, NW * 32, smem, stream>>>(p);
  CGPT_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

// query-tile height: 2 warps for <= 32 query rows (Q-Former), 5 warps when 64 < Tq <= 80 (the 72-row
// Llama prefill fits ONE tile instead of 64 + 8); otherwise 4 or 6 warps, whichever pads Tq less
// (T=257: 3 tiles of 96 = 288 rows instead of 5 tiles of 64 = 320, and 3 instead of 5 K/V sweeps)
template <int HD>
static int launch_flash(const AttnParams& p, int B, int H, cudaStream_t stream) {
  if (p.Tq <= 32) return launch_flash_nw<HD, 2>(p, B, H, stream);
  if (p.Tq > 64 && p.Tq <= 80) return launch_flash_nw<HD, 5>(p, B, H, stream);
  const int pad4 = (p.Tq + 63) / 64 * 64, pad6 = (p.Tq + 95) / 96 * 96;
  if (pad6 < pad4) return launch_flash_nw<HD, 6>(p, B, H, stream);
  return launch_flash_nw<HD, 4>(p, B, H, stream);
}

int attention(const cgpt_attn_args* a, cudaStream_t stream) {
  CGPT_REQUIRE(a != nullptr, "attention: null args");
  CGPT_REQUIRE(a->B > 0 && a->B <= 65535 && a->H > 0 && a->Tq > 0 && a->Tk > 0,
               "attention: bad sizes B=%d H=%d Tq=%d Tk=%d", a->B, a->H, a->Tq, a->Tk);
  CGPT_REQUIRE(a->head_dim % 8 == 0 && a->head_dim <= 128, "attention: head_dim %d unsupported", a->head_dim);
  if (a->head_major)   // head-major q / k / v: the pipelined tcgen05 kernels (<= 256 (+1) keys: one-tile, else multi-tile) or an error
    return attn_vit_supported(a) ? attention_vit(a, stream) : attention_long(a, stream);
  CGPT_REQUIRE(a->ldq % 8 == 0 && a->ldk % 8 == 0 && a->ldv % 8 == 0 && a->ldo % 2 == 0,
               "attention: leading dims must keep 16-byte row alignment");
  CGPT_REQUIRE(a->P >= 0 && a->P <= a->Tk && (a->P == 0 || (a->kp && a->vp)), "attention: bad prefix");
  CGPT_REQUIRE(a->Tk - a->P <= a->kv_rows_per_batch && a->Tq <= a->q_rows_per_batch,
               "attention: rows per batch too small");
  // tcgen05 one-shot kernel for the ViT / Llama-prefill shapes; mma.sync flash kernel otherwise
  // (CGPT_ATTN_LEGACY=1 forces the flash kernel, used by the parity tests to cover both)
  static const bool legacy = getenv("CGPT_ATTN_LEGACY") != nullptr;
  // persistent pipelined kernel for short sequences with 128-wide heads (the Llama prefill shape)
  if (!legacy && a->decode_kernel != 2 && attn_prefill_supported(a)) return attention_prefill(a, stream);
  if (!legacy && a->decode_kernel != 2 && attn_umma_supported(a)) return attention_umma(a, stream);
  AttnParams p;
  p.q = (const __nv_bfloat16*)a->q; p.ldq = a->ldq; p.q_rows_per_batch = a->q_rows_per_batch;
  p.k = (const __nv_bfloat16*)a->k; p.v = (const __nv_bfloat16*)a->v; p.ldk = a->ldk; p.ldv = a->ldv;
  p.kv_rows_per_batch = a->kv_rows_per_batch;
  p.kp = (const __nv_bfloat16*)a->kp; p.vp = (const __nv_bfloat16*)a->vp; p.ldkp = a->ldkp; p.ldvp = a->ldvp;
  p.P = a->P;
  p.o = (__nv_bfloat16*)a->o; p.ldo = a->ldo;
  p.Tq = a->Tq; p.Tk = a->Tk; p.head_dim = a->head_dim;
  p.scale_log2e = a->scale * 1.4426950408889634f;
  p.causal = a->causal;
  if (a->Tq == 1 && a->decode_kernel == 1) {
    CGPT_REQUIRE(a->H % 4 == 0 && a->Tk <= 1024, "decode attention: H %% 4 == 0 and Tk <= 1024 required");
    dim3 grid(a->H / 4, a->B);
    if (a->head_dim == 128) decode_attn_kernel<128><<<grid, 128, 0, stream>>>(p);
    else if (a->head_dim == 64) decode_attn_kernel<64><<<grid, 128, 0, stream>>>(p);
    else if (a->head_dim == 32) decode_attn_kernel<32><<<grid, 128, 0, stream>>>(p);
    else CGPT_REQUIRE(false, "decode attention: head_dim %d unsupported", a->head_dim);
    CGPT_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return 0;
  }
  if (a->head_dim <= 32) return launch_flash<32>(p, a->B, a->H, stream);
  if (a->head_dim <= 64) return launch_flash<64>(p, a->B, a->H, stream);
  if (a->head_dim <= 96) return launch_flash<96>(p, a->B, a->H, stream);
  return launch_flash<128>(p, a->B, a->H, stream);
}

}  // namespace cgpt
