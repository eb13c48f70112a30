// Persistent, pipelined tcgen05 attention for SHORT sequences (Tq, Tk <= 128, head_dim 128): the Llama
// prefill of the smoothing path (72 query rows x 79 keys per sample and head, causal, HF LlamaAttention
// under minigpt_base.py:414-427).  One (sample, head) is far too little work for a CTA of its own — the
// one-shot kernel (attn_umma.cu) spent ~17 us per CTA on setup and load latency for ~1 us of math, 5x
// above the HBM floor of the layer.  Here one CTA per SM walks its items through three decoupled
// pipelines:
//   warp 0     TMA producer : Q [q_pad x 128], K, V [nk_pad x 128] of item j+NL-1 into an NL-deep smem
//                             ring (128B-swizzled 64-column sub-tiles), as soon as PV(j-1) released a slot
//   warp 1     MMA issuer   : S(j+1) = Q K^T is issued BEFORE P V(j), so the tensor core never waits for a
//                             softmax; S and O of item j live in TMEM stage j & 1 (2 x [S 128 | O 128] columns)
//   warps 2-5, 6-9          : two softmax/epilogue groups, group g owns TMEM stage g: thread = query row,
//                             row max and exp2 straight out of TMEM, unnormalised P as bf16 into its own smem
//                             tile, then O / rowsum -> bf16 -> global while the other group runs its softmax
// Numerics are those of attn_umma.cu (same exp2/fma formulation, bf16 P, fp32 row sums).
#include <stdlib.h>
#include "common.cuh"
#include "ops.h"

namespace cgpt {

struct PrefillAttnParams {
  __nv_bfloat16* o; long long ldo;
  int q_rows_per_batch, kv_rows_per_batch;
  int H, Tq, Tk, causal;
  float scale_log2e;
  int q_pad, nk_pad;        // query rows padded to 8, keys padded to 16
  int n_items, n_load;      // B*H; depth of the smem load ring
  int qs, ks, stage_bytes;  // sub-tile strides (bytes) and bytes per load stage
  int tail_pad;             // bytes between the load ring and the barriers
};

constexpr int AP_THREADS = 320;
constexpr int AP_MAX_LOAD = 4;

__device__ __forceinline__ uint64_t ap_desc_mnmajor(uint32_t addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>(1024u >> 4) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}
__device__ __forceinline__ float ap_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(AP_THREADS, 1)
attn_prefill_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                    const __grid_constant__ CUtensorMap map_v, PrefillAttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by pointer arithmetic on the __shared__ array: a round trip through uintptr_t would
  // lose the address space and turn every shared-memory access below into a generic LD.E / ST.E
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  // [P0 | P1 | load stage 0 .. NL-1 | barriers].  The UMMA A operand always spans 128 rows: for q_pad < 128 it
  // reads past a Q / P sub-tile into the bytes that follow (finite or not, those rows are never stored), so
  // the P tiles sit in FRONT of the load ring and nothing is read beyond the allocation.
  const int p_bytes = 2 * p.qs;
  uint8_t* sP = smem;
  uint8_t* sL = smem + 2 * p_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sL + p.n_load * p.stage_bytes + p.tail_pad);
  uint64_t* full_qk = bars;                       // [NL]
  uint64_t* full_v = bars + AP_MAX_LOAD;          // [NL]
  uint64_t* load_free = bars + 2 * AP_MAX_LOAD;   // [NL]
  uint64_t* s_done = bars + 3 * AP_MAX_LOAD;      // [2]
  uint64_t* p_ready = s_done + 2;                 // [2] 128 arrivals
  uint64_t* o_done = s_done + 4;                  // [2]
  uint64_t* tmem_free = s_done + 6;               // [2] 128 arrivals
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(s_done + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_my = (p.n_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  const int NL = p.n_load;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_q); tma_prefetch_desc(&map_k); tma_prefetch_desc(&map_v);
    for (int i = 0; i < NL; ++i) { mbar_init(&full_qk[i], 1); mbar_init(&full_v[i], 1); mbar_init(&load_free[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_done[i], 1); mbar_init(&p_ready[i], 128); mbar_init(&o_done[i], 1); mbar_init(&tmem_free[i], 128);
    }
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

  if (warp == 0) {
    if (lane == 0) {
      // ---------------------------------------------------------------- TMA producer
      const uint32_t qk_tx = static_cast<uint32_t>(2 * p.qs + 2 * p.ks), v_tx = static_cast<uint32_t>(2 * p.ks);
      for (int j = 0; j < n_my; ++j) {
        const int ls = j % NL, u = j / NL;
        if (u >= 1) mbar_wait(&load_free[ls], (u - 1) & 1);
        const int item = static_cast<int>(blockIdx.x) + j * static_cast<int>(gridDim.x);
        const int b = item / p.H, hcol = (item % p.H) * 128;
        const int qrow = b * p.q_rows_per_batch, krow = b * p.kv_rows_per_batch;
        uint8_t* st = sL + ls * p.stage_bytes;
        mbar_arrive_expect_tx(&full_qk[ls], qk_tx);
        tma_load_2d(st, &map_q, &full_qk[ls], hcol, qrow);
        tma_load_2d(st + p.qs, &map_q, &full_qk[ls], hcol + 64, qrow);
        tma_load_2d(st + 2 * p.qs, &map_k, &full_qk[ls], hcol, krow);
        tma_load_2d(st + 2 * p.qs + p.ks, &map_k, &full_qk[ls], hcol + 64, krow);
        mbar_arrive_expect_tx(&full_v[ls], v_tx);
        tma_load_2d(st + 2 * p.qs + 2 * p.ks, &map_v, &full_v[ls], hcol, krow);
        tma_load_2d(st + 2 * p.qs + 3 * p.ks, &map_v, &full_v[ls], hcol + 64, krow);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---------------------------------------------------------------- MMA issuer
      const uint32_t idesc_s = make_idesc_bf16(128, p.nk_pad);
      const uint32_t idesc_o = make_idesc_bf16(128, 128) | (1u << 16);   // B (= V) is MN-major
      const int ksteps_o = p.nk_pad / 16;
      auto issue_s = [&](int j) {
        const int ls = j % NL, ts = j & 1;
        mbar_wait(&full_qk[ls], (j / NL) & 1);
        tcgen05_fence_after();
        const uint8_t* st = sL + ls * p.stage_bytes;
        const uint32_t d_tmem = tmem_base + ts * 256;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint64_t ad = make_smem_desc_sw128(smem_u32(st + (ks >> 2) * p.qs)) + 2 * (ks & 3);
          const uint64_t bd = make_smem_desc_sw128(smem_u32(st + 2 * p.qs + (ks >> 2) * p.ks)) + 2 * (ks & 3);
          umma_bf16(d_tmem, ad, bd, idesc_s, ks != 0);
        }
        umma_commit(&s_done[ts]);
      };
      if (n_my > 0) issue_s(0);
      for (int j = 0; j < n_my; ++j) {
        // S of the NEXT item first: its TMEM columns were released when PV(j-1) waited for p_ready(j-1)
        // (with a single load stage item j+1 can only be loaded after PV(j): keep program order then)
        if (NL >= 2 && j + 1 < n_my) issue_s(j + 1);
        const int ls = j % NL, ts = j & 1, k = j >> 1;
        mbar_wait(&full_v[ls], (j / NL) & 1);
        mbar_wait(&p_ready[ts], k & 1);
        if (k >= 1) mbar_wait(&tmem_free[ts], (k - 1) & 1);   // epilogue(j-2) has drained O of this stage
        tcgen05_fence_after();
        const uint8_t* sv = sL + ls * p.stage_bytes + 2 * p.qs + 2 * p.ks;
        const uint8_t* sp = sP + ts * p_bytes;
        const uint32_t d_tmem = tmem_base + ts * 256 + 128;
        for (int ks = 0; ks < ksteps_o; ++ks) {
          const uint64_t ad = make_smem_desc_sw128(smem_u32(sp + (ks >> 2) * p.qs)) + 2 * (ks & 3);
          const uint64_t bd = ap_desc_mnmajor(smem_u32(sv + ks * 16 * 128), static_cast<uint32_t>(p.ks));
          umma_bf16(d_tmem, ad, bd, idesc_o, ks != 0);
        }
        umma_commit(&o_done[ts]);
        umma_commit(&load_free[ls]);
        if (NL < 2 && j + 1 < n_my) issue_s(j + 1);
      }
      // do not leave with an mbarrier arrive still in flight
      if (n_my > 0) mbar_wait(&load_free[(n_my - 1) % NL], ((n_my - 1) / NL) & 1);
    }
  } else {
    // ---------------------------------------------------------------- softmax + epilogue, thread = query row
    const int g = (warp - 2) >> 2;                     // group <-> TMEM / P stage
    const int r = (warp & 3) * 32 + lane;              // TMEM lane quarter of this warp is warp % 4
    const bool row_ok = r < p.Tq;
    const int offs = p.Tk - p.Tq;
    const int kmax = p.causal ? min(p.Tk, r + offs + 1) : p.Tk;   // keys [0, kmax) visible
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + g * 256;
    uint8_t* sp = sP + g * p_bytes;
    for (int j = g; j < n_my; j += 2) {
      const int k = j >> 1;
      const int item = static_cast<int>(blockIdx.x) + j * static_cast<int>(gridDim.x);
      const int b = item / p.H, hcol = (item % p.H) * 128;
      mbar_wait(&s_done[g], k & 1);
      tcgen05_fence_after();
      // pass 1: row max of the raw scores
      float mx = -INFINITY;
      {
        uint32_t v[16];
#pragma unroll 1
        for (int c = 0; c < p.nk_pad; c += 16) {
          tmem_ld_x16(t_lane + c, v);
          tmem_ld_wait();
          if (c + 16 <= kmax) {
#pragma unroll
            for (int i = 0; i < 16; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) if (c + i < kmax) mx = fmaxf(mx, __uint_as_float(v[i]));
          }
        }
      }
      if (mx == -INFINITY) mx = 0.f;
      const float neg_ms = -mx * p.scale_log2e;
      // pass 2: exp2, row sum, unnormalised P (bf16, K-major, 128B swizzle) into this group's tile
      float sum = 0.f;
      {
        uint32_t v[16];
#pragma unroll 1
        for (int c = 0; c < p.nk_pad; c += 16) {
          tmem_ld_x16(t_lane + c, v);
          tmem_ld_wait();
          float e[16];
          if (c + 16 <= kmax) {
#pragma unroll
            for (int i = 0; i < 16; ++i) e[i] = ap_exp2(fmaf(__uint_as_float(v[i]), p.scale_log2e, neg_ms));
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              e[i] = (c + i < kmax) ? ap_exp2(fmaf(__uint_as_float(v[i]), p.scale_log2e, neg_ms)) : 0.f;
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) sum += e[i];
          if (r < p.q_pad) {
            const int st = c >> 6, j0 = (c & 63) >> 3;
            uint8_t* rowp = sp + st * p.qs + r * 128;
            *reinterpret_cast<uint4*>(rowp + (((j0) ^ (r & 7)) << 4)) =
                make_uint4(pack_bf16x2(e[0], e[1]), pack_bf16x2(e[2], e[3]), pack_bf16x2(e[4], e[5]), pack_bf16x2(e[6], e[7]));
            *reinterpret_cast<uint4*>(rowp + (((j0 + 1) ^ (r & 7)) << 4)) =
                make_uint4(pack_bf16x2(e[8], e[9]), pack_bf16x2(e[10], e[11]), pack_bf16x2(e[12], e[13]), pack_bf16x2(e[14], e[15]));
          }
        }
      }
      fence_proxy_async();        // P (generic-proxy stores) -> visible to the tensor core
      tcgen05_fence_before();     // this thread's TMEM reads of S are done before S(j+2) may overwrite them
      mbar_arrive(&p_ready[g]);
      // ---- epilogue: O / rowsum -> bf16 -> global
      mbar_wait(&o_done[g], k & 1);
      tcgen05_fence_after();
      const float inv = sum > 0.f ? 1.f / sum : 0.f;
      __nv_bfloat16* orow = p.o + (static_cast<long long>(b) * p.q_rows_per_batch + r) * p.ldo + hcol;
      {
        uint32_t va[16], vb[16];
        tmem_ld_x16(t_lane + 128, va);
#pragma unroll 1
        for (int c = 0; c < 128; c += 32) {
          tmem_ld_wait();
          tmem_ld_x16(t_lane + 128 + c + 16, vb);
          if (row_ok) {
            *reinterpret_cast<uint4*>(orow + c) =
                make_uint4(pack_bf16x2(__uint_as_float(va[0]) * inv, __uint_as_float(va[1]) * inv),
                           pack_bf16x2(__uint_as_float(va[2]) * inv, __uint_as_float(va[3]) * inv),
                           pack_bf16x2(__uint_as_float(va[4]) * inv, __uint_as_float(va[5]) * inv),
                           pack_bf16x2(__uint_as_float(va[6]) * inv, __uint_as_float(va[7]) * inv));
            *reinterpret_cast<uint4*>(orow + c + 8) =
                make_uint4(pack_bf16x2(__uint_as_float(va[8]) * inv, __uint_as_float(va[9]) * inv),
                           pack_bf16x2(__uint_as_float(va[10]) * inv, __uint_as_float(va[11]) * inv),
                           pack_bf16x2(__uint_as_float(va[12]) * inv, __uint_as_float(va[13]) * inv),
                           pack_bf16x2(__uint_as_float(va[14]) * inv, __uint_as_float(va[15]) * inv));
          }
          tmem_ld_wait();
          if (c + 32 < 128) tmem_ld_x16(t_lane + 128 + c + 32, va);
          if (row_ok) {
            *reinterpret_cast<uint4*>(orow + c + 16) =
                make_uint4(pack_bf16x2(__uint_as_float(vb[0]) * inv, __uint_as_float(vb[1]) * inv),
                           pack_bf16x2(__uint_as_float(vb[2]) * inv, __uint_as_float(vb[3]) * inv),
                           pack_bf16x2(__uint_as_float(vb[4]) * inv, __uint_as_float(vb[5]) * inv),
                           pack_bf16x2(__uint_as_float(vb[6]) * inv, __uint_as_float(vb[7]) * inv));
            *reinterpret_cast<uint4*>(orow + c + 24) =
                make_uint4(pack_bf16x2(__uint_as_float(vb[8]) * inv, __uint_as_float(vb[9]) * inv),
                           pack_bf16x2(__uint_as_float(vb[10]) * inv, __uint_as_float(vb[11]) * inv),
                           pack_bf16x2(__uint_as_float(vb[12]) * inv, __uint_as_float(vb[13]) * inv),
                           pack_bf16x2(__uint_as_float(vb[14]) * inv, __uint_as_float(vb[15]) * inv));
          }
        }
      }
      tcgen05_fence_before();     // O reads done before PV(j+2) may overwrite this stage
      mbar_arrive(&tmem_free[g]);
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
typedef CUresult (*EncodeTiledFn3)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn3 ap_encode_fn() {
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
static int ap_make_map(CUtensorMap* map, const void* base, long long rows, long long cols, long long ld, int box_rows) {
  EncodeTiledFn3 fn = ap_encode_fn();
  CGPT_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled driver entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CGPT_REQUIRE(r == CUDA_SUCCESS, "attention_prefill: cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld box=64x%d",
               (int)r, rows, cols, ld, box_rows);
  return 0;
}

// 1 = this shape is served by the persistent short-sequence kernel
int attn_prefill_supported(const cgpt_attn_args* a) {
  static const bool off = getenv("CGPT_ATTN_NO_PREFILL") != nullptr;
  if (off || a->decode_kernel != 0 || a->P != 0) return 0;
  if (a->head_dim != 128 || a->Tq > 128 || a->Tk > 128 || a->Tq < 1 || a->Tk < 1) return 0;
  if (a->causal && a->Tk < a->Tq) return 0;
  if (a->B * (long long)a->H > 0x7fffffffLL) return 0;
  if ((reinterpret_cast<uintptr_t>(a->q) | reinterpret_cast<uintptr_t>(a->k) | reinterpret_cast<uintptr_t>(a->v)) & 15) return 0;
  if ((a->ldq | a->ldk | a->ldv | a->ldo) & 7) return 0;
  return 1;
}

int attention_prefill(const cgpt_attn_args* a, cudaStream_t stream) {
  PrefillAttnParams p;
  p.o = (__nv_bfloat16*)a->o; p.ldo = a->ldo;
  p.q_rows_per_batch = a->q_rows_per_batch; p.kv_rows_per_batch = a->kv_rows_per_batch;
  p.H = a->H; p.Tq = a->Tq; p.Tk = a->Tk; p.causal = a->causal;
  p.scale_log2e = a->scale * 1.4426950408889634f;
  p.q_pad = (a->Tq + 7) / 8 * 8;
  p.nk_pad = (a->Tk + 15) / 16 * 16;
  p.n_items = a->B * a->H;
  p.qs = p.q_pad * 128;
  p.ks = p.nk_pad * 128;
  p.stage_bytes = 2 * p.qs + 4 * p.ks;
  // the UMMA reads the second Q sub-tile of the last load stage 128 rows deep: pad the ring's tail so that
  // stays inside the allocation when the stage itself is shorter (tiny shapes only)
  const int over = p.qs + 128 * 128 - p.stage_bytes;
  p.tail_pad = over > 0 ? (over + 1023) / 1024 * 1024 : 0;
  const int fixed = 2 * (2 * p.qs) + p.tail_pad + (3 * AP_MAX_LOAD + 8) * 8 + 16 + 1024;
  int n_load = (227 * 1024 - fixed) / p.stage_bytes;
  if (n_load > AP_MAX_LOAD) n_load = AP_MAX_LOAD;
  CGPT_REQUIRE(n_load >= 1, "attention_prefill: shared memory too small for Tq=%d Tk=%d", a->Tq, a->Tk);
  p.n_load = n_load;
  const int smem = fixed + n_load * p.stage_bytes;

  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    CGPT_CHECK_CUDA(cudaGetDevice(&dev));
    CGPT_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const long long q_rows = (long long)a->B * a->q_rows_per_batch, kv_rows = (long long)a->B * a->kv_rows_per_batch;
  const long long cols = (long long)a->H * 128;
  CUtensorMap mq, mk, mv;
  if (int rc = ap_make_map(&mq, a->q, q_rows, cols, a->ldq, p.q_pad)) return rc;
  if (int rc = ap_make_map(&mk, a->k, kv_rows, cols, a->ldk, p.nk_pad)) return rc;
  if (int rc = ap_make_map(&mv, a->v, kv_rows, cols, a->ldv, p.nk_pad)) return rc;
  static int configured_smem = 0;
  if (smem > configured_smem) {
    CGPT_CHECK_CUDA(cudaFuncSetAttribute(attn_prefill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured_smem = smem;
  }
  const int grid = p.n_items < sms ? p.n_items : sms;
  attn_prefill_kernel<<<grid, AP_THREADS, smem, stream>>>(mq, mk, mv, p);
  CGPT_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace cgpt
