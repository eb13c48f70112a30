// Backward / optimiser kernels of the noise-augmented fine-tune step:
// MiniGPT4FineTuneAgent.train (agents/minigpt4_finetune_agent.py:149-195) does loss.backward() through the frozen
// Llama to the one trainable module, llama_proj (base_model.py:162-172,238-240; minigpt4.py:76-78,111-117), and an
// AdamW step.  The data-gradient GEMMs reuse the tcgen05 GEMM on transposed weight copies; everything that is not a
// GEMM lives here.  Training batches are a handful of images (a few hundred token rows), so these kernels are
// written for clarity and exactness (fp32 math, fixed summation order), not for the roofline.
#include <math.h>
#include "common.cuh"
#include "ops.h"

namespace cgpt {
namespace {

__device__ __forceinline__ float bf(const __nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// ---------------------------------------------------------------- SwiGLU (HF LlamaMLP: down(silu(gate) * up))
// gu: [M, 2I] bf16 with (gate_j, up_j) interleaved (the packing of the fused gate/up weight)
__global__ void swiglu_fwd_kernel(const __nv_bfloat16* __restrict__ gu, __nv_bfloat16* __restrict__ act, long long n) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const float g = bf(gu[2 * i]), u = bf(gu[2 * i + 1]);
  act[i] = __float2bfloat16(g * sigmoidf_(g) * u);
}
__global__ void swiglu_bwd_kernel(const __nv_bfloat16* __restrict__ gu, const __nv_bfloat16* __restrict__ dact,
                                  __nv_bfloat16* __restrict__ dgu, long long n) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const float g = bf(gu[2 * i]), u = bf(gu[2 * i + 1]), d = bf(dact[i]);
  const float s = sigmoidf_(g);
  dgu[2 * i] = __float2bfloat16(d * u * s * (1.0f + g * (1.0f - s)));
  dgu[2 * i + 1] = __float2bfloat16(d * g * s);
}

// ---------------------------------------------------------------- RMSNorm backward (HF LlamaRMSNorm)
// h = x * r * gamma, r = rsqrt(mean(x^2) + eps).  dx[row] += r * gamma * dy - x * r^3 * mean(x * gamma * dy).
// Row r of dy maps to row (r / period) * stride + offset + r % period of x / dx when period > 0 (the final norm is
// applied to the answer-predicting rows only).  One block per row, fixed-order reduction.
__global__ void __launch_bounds__(256) rmsnorm_bwd_kernel(const float* __restrict__ x, long long ldx,
                                                          const float* __restrict__ gamma, const float* __restrict__ dy,
                                                          long long ldy, float eps, int D, float* __restrict__ dx,
                                                          long long lddx, int period, int stride, int offset) {
  __shared__ float red[2][8];
  const int r = blockIdx.x;
  const long long xr = period > 0 ? static_cast<long long>(r / period) * stride + offset + r % period : r;
  const float* xp = x + xr * ldx;
  const float* dyp = dy + r * ldy;
  float ss = 0.f, sd = 0.f;
  for (int c = threadIdx.x; c < D; c += 256) {
    const float xv = xp[c];
    ss = fmaf(xv, xv, ss);
    sd = fmaf(xv * gamma[c], dyp[c], sd);
  }
  ss = warp_sum(ss); sd = warp_sum(sd);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = ss; red[1][threadIdx.x >> 5] = sd; }
  __syncthreads();
  float tss = 0.f, tsd = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { tss += red[0][i]; tsd += red[1][i]; }
  const float rstd = rsqrtf(tss / D + eps);
  const float coef = rstd * rstd * rstd * tsd / D;
  float* dxp = dx + xr * lddx;
  for (int c = threadIdx.x; c < D; c += 256) dxp[c] += rstd * gamma[c] * dyp[c] - xp[c] * coef;
}

// ---------------------------------------------------------------- rotary backward + cast
// dqkv: f32 [M, 3*H*hd] (gradients wrt the ROTATED q, k and wrt v) -> bf16 [M, 3*H*hd] gradients wrt the q/k/v
// projections: the transpose of HF's rotate_half rotation on the q and k parts (angle negated), v copied.
__global__ void rope_bwd_cast_kernel(const float* __restrict__ dqkv, __nv_bfloat16* __restrict__ out, int T, int H, int hd,
                                     int pos0, const float* __restrict__ cos_t, const float* __restrict__ sin_t) {
  const int m = blockIdx.x;
  const int pos = pos0 + m % T;
  const int half = hd / 2, D = H * hd;
  const float* src = dqkv + static_cast<long long>(m) * 3 * D;
  __nv_bfloat16* dst = out + static_cast<long long>(m) * 3 * D;
  const float* cs = cos_t + static_cast<long long>(pos) * half;
  const float* sn = sin_t + static_cast<long long>(pos) * half;
  for (int w = threadIdx.x; w < 2 * H * half; w += blockDim.x) {
    const int part = w / (H * half), rem = w - part * H * half;
    const int h = rem / half, j = rem - h * half;
    const int base = part * D + h * hd;
    const float g1 = src[base + j], g2 = src[base + half + j];
    // forward: o1 = x1 c - x2 s ; o2 = x2 c + x1 s   =>   dx1 = g1 c + g2 s ; dx2 = g2 c - g1 s
    dst[base + j] = __float2bfloat16(g1 * cs[j] + g2 * sn[j]);
    dst[base + half + j] = __float2bfloat16(g2 * cs[j] - g1 * sn[j]);
  }
  for (int c = threadIdx.x; c < D; c += blockDim.x) dst[2 * D + c] = __float2bfloat16(src[2 * D + c]);
}

// ---------------------------------------------------------------- causal attention backward (short sequences)
// One CTA per (sample, head).  q: rotated queries [B*Tq, ldq]; K / V: cache rows [0, Tk) of the sample (prefix keys
// included), Tk - Tq = number of keys before the first query; o: forward output; dout: gradient wrt o.
// dqkv (f32 [B*Tq, 3*H*hd]) receives dq, and dk / dv for the keys that belong to this sample's own rows
// (key j >= Tk - Tq <-> row j - (Tk - Tq)); gradients of the shared prefix keys are dropped (no parameter behind them).
// Q, K, V, dO sit in shared memory as bf16 with rows padded by 16 bytes (consecutive rows start 4 banks apart, so a
// warp reading 32 different rows at one column offset needs the minimum 4 wavefronts); every inner loop consumes
// 8 bf16 per 16-byte load.
__device__ __forceinline__ float dot8(const uint4 a, const uint4 b) {
  return bf16_lo(a.x) * bf16_lo(b.x) + bf16_hi(a.x) * bf16_hi(b.x) + bf16_lo(a.y) * bf16_lo(b.y) + bf16_hi(a.y) * bf16_hi(b.y) +
         bf16_lo(a.z) * bf16_lo(b.z) + bf16_hi(a.z) * bf16_hi(b.z) + bf16_lo(a.w) * bf16_lo(b.w) + bf16_hi(a.w) * bf16_hi(b.w);
}
__device__ __forceinline__ void axpy8(float* acc, float w, const uint4 v) {
  acc[0] = fmaf(w, bf16_lo(v.x), acc[0]); acc[1] = fmaf(w, bf16_hi(v.x), acc[1]);
  acc[2] = fmaf(w, bf16_lo(v.y), acc[2]); acc[3] = fmaf(w, bf16_hi(v.y), acc[3]);
  acc[4] = fmaf(w, bf16_lo(v.z), acc[4]); acc[5] = fmaf(w, bf16_hi(v.z), acc[5]);
  acc[6] = fmaf(w, bf16_lo(v.w), acc[6]); acc[7] = fmaf(w, bf16_hi(v.w), acc[7]);
}
__global__ void __launch_bounds__(256) attn_bwd_kernel(const __nv_bfloat16* __restrict__ q, long long ldq,
                                                       const __nv_bfloat16* __restrict__ kc,
                                                       const __nv_bfloat16* __restrict__ vc, long long ldc, int cache_rows,
                                                       const __nv_bfloat16* __restrict__ o, long long ldo,
                                                       const __nv_bfloat16* __restrict__ dout, long long lddo,
                                                       float* __restrict__ dqkv, int H, int hd, int Tq, int Tk, float scale) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int hp = hd + 8;                                                   // padded row, elements
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem_raw);          // [Tq, hp]
  __nv_bfloat16* sK = sQ + Tq * hp;                                        // [Tk, hp]
  __nv_bfloat16* sV = sK + Tk * hp;                                        // [Tk, hp]
  __nv_bfloat16* sD = sV + Tk * hp;                                        // [Tq, hp]  dO
  float* sP = reinterpret_cast<float*>(sD + Tq * hp);                      // [Tq, Tk]  P, then dS
  float* sDelta = sP + Tq * Tk;                                            // [Tq]
  const int h = blockIdx.x % H, b = blockIdx.x / H;
  const int off = Tk - Tq, D = H * hd, nc = hd >> 3;                       // nc: 16-byte chunks per head row
  const long long qrow0 = static_cast<long long>(b) * Tq, krow0 = static_cast<long long>(b) * cache_rows;
  auto row16 = [&](const __nv_bfloat16* base, int r, int c) { return *reinterpret_cast<const uint4*>(base + r * hp + c * 8); };
  for (int e = threadIdx.x; e < Tq * nc; e += 256) {
    const int r = e / nc, c = e - r * nc;
    *reinterpret_cast<uint4*>(sQ + r * hp + c * 8) = *reinterpret_cast<const uint4*>(q + (qrow0 + r) * ldq + h * hd + c * 8);
    *reinterpret_cast<uint4*>(sD + r * hp + c * 8) = *reinterpret_cast<const uint4*>(dout + (qrow0 + r) * lddo + h * hd + c * 8);
  }
  for (int e = threadIdx.x; e < Tk * nc; e += 256) {
    const int r = e / nc, c = e - r * nc;
    *reinterpret_cast<uint4*>(sK + r * hp + c * 8) = *reinterpret_cast<const uint4*>(kc + (krow0 + r) * ldc + h * hd + c * 8);
    *reinterpret_cast<uint4*>(sV + r * hp + c * 8) = *reinterpret_cast<const uint4*>(vc + (krow0 + r) * ldc + h * hd + c * 8);
  }
  __syncthreads();
  // S = scale * Q K^T with the causal mask (query i sees keys <= i + off)
  for (int e = threadIdx.x; e < Tq * Tk; e += 256) {
    const int i = e / Tk, j = e - i * Tk;
    float acc = -INFINITY;
    if (j <= i + off) {
      acc = 0.f;
      for (int c = 0; c < nc; ++c) acc += dot8(row16(sQ, i, c), row16(sK, j, c));
      acc *= scale;
    }
    sP[e] = acc;
  }
  __syncthreads();
  // row softmax and delta_i = dO_i . O_i (one warp per row, fixed order)
  for (int i = threadIdx.x >> 5; i < Tq; i += 8) {
    const int lane = threadIdx.x & 31;
    float mx = -INFINITY;
    for (int j = lane; j < Tk; j += 32) mx = fmaxf(mx, sP[i * Tk + j]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < Tk; j += 32) {
      const float e = expf(sP[i * Tk + j] - mx);
      sP[i * Tk + j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    for (int j = lane; j < Tk; j += 32) sP[i * Tk + j] *= inv;
    float dl = 0.f;
    for (int c = lane; c < hd; c += 32) dl = fmaf(bf(sD[i * hp + c]), bf(o[(qrow0 + i) * ldo + h * hd + c]), dl);
    dl = warp_sum(dl);
    if (lane == 0) sDelta[i] = dl;
  }
  __syncthreads();
  // dV[j, 8 cols] = sum_i P_ij dO_i  (own keys only; P is zero above the diagonal, so i starts at the key's own row)
  for (int e = threadIdx.x; e < Tq * nc; e += 256) {
    const int jr = e / nc, c = e - jr * nc;            // own row jr <-> key j = jr + off
    const int j = jr + off;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int i = jr; i < Tq; ++i) axpy8(acc, sP[i * Tk + j], row16(sD, i, c));
    float* dst = dqkv + (qrow0 + jr) * 3 * D + 2 * D + h * hd + c * 8;
    *reinterpret_cast<float4*>(dst) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    *reinterpret_cast<float4*>(dst + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
  }
  __syncthreads();
  // dS_ij = P_ij (dO_i . V_j - delta_i), in place
  for (int e = threadIdx.x; e < Tq * Tk; e += 256) {
    const int i = e / Tk, j = e - i * Tk;
    float v = 0.f;
    if (j <= i + off) {
      float dp = 0.f;
      for (int c = 0; c < nc; ++c) dp += dot8(row16(sD, i, c), row16(sV, j, c));
      v = sP[e] * (dp - sDelta[i]);
    }
    sP[e] = v;
  }
  __syncthreads();
  // dQ_i = scale * sum_j dS_ij K_j ;  dK_j = scale * sum_i dS_ij Q_i (own keys only)
  for (int e = threadIdx.x; e < Tq * nc; e += 256) {
    const int i = e / nc, c = e - i * nc;
    float aq[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int j = 0; j <= i + off; ++j) axpy8(aq, sP[i * Tk + j], row16(sK, j, c));
    float* dq = dqkv + (qrow0 + i) * 3 * D + h * hd + c * 8;
    *reinterpret_cast<float4*>(dq) = make_float4(aq[0] * scale, aq[1] * scale, aq[2] * scale, aq[3] * scale);
    *reinterpret_cast<float4*>(dq + 4) = make_float4(aq[4] * scale, aq[5] * scale, aq[6] * scale, aq[7] * scale);
    const int j = i + off;                             // own key of row i
    float ak[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int ii = i; ii < Tq; ++ii) axpy8(ak, sP[ii * Tk + j], row16(sQ, ii, c));
    float* dk = dqkv + (qrow0 + i) * 3 * D + D + h * hd + c * 8;
    *reinterpret_cast<float4*>(dk) = make_float4(ak[0] * scale, ak[1] * scale, ak[2] * scale, ak[3] * scale);
    *reinterpret_cast<float4*>(dk + 4) = make_float4(ak[4] * scale, ak[5] * scale, ak[6] * scale, ak[7] * scale);
  }
}

// ---------------------------------------------------------------- cross-entropy gradient
// dlogits[r, :] = (softmax(logits[r, :]) - (1 - eps) * onehot(target[r]) - eps / cols) / count ; rows with target < 0 get
// zeros (eps = label smoothing, modeling_llama.py:107)
__global__ void __launch_bounds__(256) ce_grad_kernel(const float* __restrict__ logits, long long ld, int cols,
                                                      const int* __restrict__ targets, const float* __restrict__ mean_count,
                                                      __nv_bfloat16* __restrict__ dlogits, long long ldd, float eps) {
  __shared__ float red[8];
  const int r = blockIdx.x;
  const int t = targets[r];
  __nv_bfloat16* out = dlogits + r * ldd;
  if (t < 0 || t >= cols) {
    for (int c = threadIdx.x; c < cols; c += 256) out[c] = __float2bfloat16(0.f);
    return;
  }
  const float* row = logits + r * ld;
  float mx = -INFINITY;
  for (int c = threadIdx.x; c < cols; c += 256) mx = fmaxf(mx, row[c]);
  mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = red[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) mx = fmaxf(mx, red[i]);
  __syncthreads();
  float sum = 0.f;
  for (int c = threadIdx.x; c < cols; c += 256) sum += expf(row[c] - mx);
  sum = warp_sum(sum);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += red[i];
  const float inv = 1.f / (tot * mean_count[1]);
  const float sub = (1.f - eps) / mean_count[1];
  const float uni = eps / (static_cast<float>(cols) * mean_count[1]);
  for (int c = threadIdx.x; c < cols; c += 256)
    out[c] = __float2bfloat16(expf(row[c] - mx) * inv - (c == t ? sub : 0.f) - uni);
}

// ---------------------------------------------------------------- small helpers
__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, long long lds, __nv_bfloat16* __restrict__ dst,
                                     long long ldd, int rows, int cols, int period, int stride, int offset) {
  // dst[r, :] = bf16(src[map(r), :]) with the optional row gather map(r) = (r / period) * stride + offset + r % period
  const int r = blockIdx.x;
  const long long sr = period > 0 ? static_cast<long long>(r / period) * stride + offset + r % period : r;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) dst[r * ldd + c] = __float2bfloat16(src[sr * lds + c]);
}
__global__ void transpose_bf16_kernel(const __nv_bfloat16* __restrict__ src, long long lds, __nv_bfloat16* __restrict__ dst,
                                      long long ldd, int rows, int cols) {
  __shared__ __nv_bfloat16 tile[32][33];
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? src[r * lds + c] : __float2bfloat16(0.f);
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + threadIdx.x;          // dst[c, r]
    if (c < cols && r < rows) dst[c * ldd + r] = tile[threadIdx.x][i];
  }
}
// out[c] = sum_r bf16(src[r, c]) in row order (bias gradient), src bf16
__global__ void colsum_bf16_kernel(const __nv_bfloat16* __restrict__ src, long long lds, int rows, int cols,
                                   float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  float acc = 0.f;
  for (int r = 0; r < rows; ++r) acc += bf(src[r * lds + c]);
  out[c] = acc;
}
// torch.optim.AdamW (decoupled weight decay), fp32 master weights + bf16 copy for the GEMMs
__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                             __nv_bfloat16* __restrict__ p_bf16, long long n, float lr, float beta1, float beta2, float eps,
                             float wd, float bc1, float bc2, float gscale) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const float gi = g[i] * gscale;
  float pi = p[i] * (1.0f - lr * wd);
  const float mi = beta1 * m[i] + (1.0f - beta1) * gi;
  const float vi = beta2 * v[i] + (1.0f - beta2) * gi * gi;
  m[i] = mi; v[i] = vi;
  pi -= lr * (mi / bc1) / (sqrtf(vi / bc2) + eps);
  p[i] = pi;
  if (p_bf16) p_bf16[i] = __float2bfloat16(pi);
}

}  // namespace

#define LAUNCH_OK()                        \
  do {                                     \
    CGPT_CHECK_CUDA(cudaGetLastError());   \
    count_launch();                        \
    return 0;                              \
  } while (0)

int swiglu_fwd(const void* gu, void* act, long long rows, int inter, cudaStream_t s) {
  const long long n = rows * inter;
  CGPT_REQUIRE(gu && act && n > 0, "swiglu_fwd: bad arguments");
  swiglu_fwd_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(static_cast<const __nv_bfloat16*>(gu),
                                                                          static_cast<__nv_bfloat16*>(act), n);
  LAUNCH_OK();
}
int swiglu_bwd(const void* gu, const void* dact, void* dgu, long long rows, int inter, cudaStream_t s) {
  const long long n = rows * inter;
  CGPT_REQUIRE(gu && dact && dgu && n > 0, "swiglu_bwd: bad arguments");
  swiglu_bwd_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(
      static_cast<const __nv_bfloat16*>(gu), static_cast<const __nv_bfloat16*>(dact), static_cast<__nv_bfloat16*>(dgu), n);
  LAUNCH_OK();
}
int rmsnorm_bwd(const float* x, long long ldx, const float* gamma, const float* dy, long long ldy, float eps, int rows, int D,
                float* dx, long long lddx, int period, int stride, int offset, cudaStream_t s) {
  CGPT_REQUIRE(x && gamma && dy && dx && rows > 0 && D > 0, "rmsnorm_bwd: bad arguments");
  rmsnorm_bwd_kernel<<<rows, 256, 0, s>>>(x, ldx, gamma, dy, ldy, eps, D, dx, lddx, period, stride, offset);
  LAUNCH_OK();
}
int rope_bwd_cast(const float* dqkv, void* out, int rows, int T, int H, int hd, int pos0, const float* cos_t,
                  const float* sin_t, cudaStream_t s) {
  CGPT_REQUIRE(dqkv && out && rows > 0 && T > 0 && hd % 2 == 0, "rope_bwd_cast: bad arguments");
  rope_bwd_cast_kernel<<<rows, 256, 0, s>>>(dqkv, static_cast<__nv_bfloat16*>(out), T, H, hd, pos0, cos_t, sin_t);
  LAUNCH_OK();
}
int attention_bwd(const void* q, long long ldq, const void* kc, const void* vc, long long ldc, int cache_rows, const void* o,
                  long long ldo, const void* dout, long long lddo, float* dqkv, int B, int H, int hd, int Tq, int Tk,
                  float scale, cudaStream_t s) {
  CGPT_REQUIRE(q && kc && vc && o && dout && dqkv, "attention_bwd: null argument");
  CGPT_REQUIRE(B > 0 && H > 0 && hd > 0 && Tq > 0 && Tk >= Tq && Tk <= cache_rows, "attention_bwd: bad sizes Tq=%d Tk=%d", Tq, Tk);
  CGPT_REQUIRE(hd % 8 == 0 && ldq % 8 == 0 && ldc % 8 == 0 && lddo % 8 == 0 && (H * hd) % 4 == 0,
               "attention_bwd: head_dim and leading dimensions must be multiples of 8");
  const size_t smem = static_cast<size_t>(2 * Tq + 2 * Tk) * (hd + 8) * 2 + static_cast<size_t>(Tq) * Tk * 4 + Tq * 4;
  CGPT_REQUIRE(smem <= 227 * 1024, "attention_bwd: Tq=%d Tk=%d hd=%d needs %zu bytes of shared memory", Tq, Tk, hd, smem);
  static size_t configured = 0;
  if (smem > configured) {
    CGPT_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    configured = smem;
  }
  attn_bwd_kernel<<<B * H, 256, smem, s>>>(static_cast<const __nv_bfloat16*>(q), ldq, static_cast<const __nv_bfloat16*>(kc),
                                           static_cast<const __nv_bfloat16*>(vc), ldc, cache_rows,
                                           static_cast<const __nv_bfloat16*>(o), ldo,
                                           static_cast<const __nv_bfloat16*>(dout), lddo, dqkv, H, hd, Tq, Tk, scale);
  LAUNCH_OK();
}
int ce_grad(const float* logits, long long ld, int rows, int cols, const int* targets, const float* mean_count, void* dlogits,
            long long ldd, float label_smoothing, cudaStream_t s) {
  CGPT_REQUIRE(logits && targets && mean_count && dlogits && rows > 0 && cols > 0, "ce_grad: bad arguments");
  CGPT_REQUIRE(label_smoothing >= 0.f && label_smoothing < 1.f, "ce_grad: label_smoothing %f outside [0, 1)", label_smoothing);
  ce_grad_kernel<<<rows, 256, 0, s>>>(logits, ld, cols, targets, mean_count, static_cast<__nv_bfloat16*>(dlogits), ldd,
                                      label_smoothing);
  LAUNCH_OK();
}
int cast_rows_f32_bf16(const float* src, long long lds, void* dst, long long ldd, int rows, int cols, int period, int stride,
                       int offset, cudaStream_t s) {
  CGPT_REQUIRE(src && dst && rows > 0 && cols > 0, "cast_rows: bad arguments");
  cast_f32_bf16_kernel<<<rows, 256, 0, s>>>(src, lds, static_cast<__nv_bfloat16*>(dst), ldd, rows, cols, period, stride, offset);
  LAUNCH_OK();
}
int transpose_bf16(const void* src, long long lds, void* dst, long long ldd, int rows, int cols, cudaStream_t s) {
  CGPT_REQUIRE(src && dst && rows > 0 && cols > 0, "transpose_bf16: bad arguments");
  transpose_bf16_kernel<<<dim3((cols + 31) / 32, (rows + 31) / 32), dim3(32, 8), 0, s>>>(
      static_cast<const __nv_bfloat16*>(src), lds, static_cast<__nv_bfloat16*>(dst), ldd, rows, cols);
  LAUNCH_OK();
}
int colsum_bf16(const void* src, long long lds, int rows, int cols, float* out, cudaStream_t s) {
  CGPT_REQUIRE(src && out && rows > 0 && cols > 0, "colsum_bf16: bad arguments");
  colsum_bf16_kernel<<<(cols + 127) / 128, 128, 0, s>>>(static_cast<const __nv_bfloat16*>(src), lds, rows, cols, out);
  LAUNCH_OK();
}
int adamw_step(float* p, const float* g, float* m, float* v, void* p_bf16, long long n, float lr, float beta1, float beta2,
               float eps, float wd, int step, float gscale, cudaStream_t s) {
  CGPT_REQUIRE(p && g && m && v && n > 0 && step >= 1, "adamw_step: bad arguments");
  const float bc1 = 1.0f - powf(beta1, static_cast<float>(step)), bc2 = 1.0f - powf(beta2, static_cast<float>(step));
  adamw_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(p, g, m, v, static_cast<__nv_bfloat16*>(p_bf16), n, lr,
                                                                     beta1, beta2, eps, wd, bc1, bc2, gscale);
  LAUNCH_OK();
}

}  // namespace cgpt
