// K1 - fused noise kernel: Philox4x32-10 Gaussian (or uniform) draw, sigma scale, add to the
// image, BLIP Normalize, and patchify to 14x14 tiles as the bf16 A-operand of the patch-embed
// GEMM, in ONE pass.
//
// Replaces, per noise batch:   batch = x.repeat(B,1,1,1); noise = randn_like(batch)*sigma;
// batch + noise                                  (randomized_smoothing/smoothing.py:95-97)
// + transforms.Normalize(mean, std)              (processors/base_processor.py:17-34)
// + the unfold implied by Conv2d(k=s=14)         (eva_vit.py:202,209)
// which in eager PyTorch is ~7 full-tensor fp32 passes (4.2 MB/sample); here 301 KB/sample
// are written once.
//
// Work decomposition: one CTA per (sample, patch-row).  Threads walk the 3 x 14 image rows of
// that patch-row in groups of 4 consecutive pixels (= one Philox counter, one float4 load of x),
// scatter the bf16 results into a shared-memory image of the (S/14) x 592 output rows, and the
// CTA then streams that contiguous block to HBM with 16-byte stores.
//
// Random stream definition (restated in oracle/philox_oracle.py):
//   key     = (seed_lo, seed_hi)
//   counter = (g, sample_lo, stream_id, sample_hi),  g = (c*S*S + y*S + x) / 4
//   the 4 outputs r0..r3 give the 4 pixels x..x+3:
//     gaussian: u = r*2^-32 + 2^-33;  (z0,z1) = sqrt(-2 ln u0) * (cos, sin)(2 pi u1), (z2,z3) from (u2,u3)
//     uniform : u = (r >> 8) * 2^-24   in [0,1)        (torch.rand_like convention, eval agent :185)
//   so the draw of sample i is independent of batch size and of the number of GPUs.
#include "common.cuh"
#include "ops.h"

namespace cgpt {

constexpr int PATCH = 14;
constexpr int PATCH_K = 3 * PATCH * PATCH;  // 588
constexpr int PATCH_K_PAD = 592;            // 16-byte rows for TMA

struct Philox4 {
  uint32_t x, y, z, w;
};

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2,
                                                          uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = static_cast<uint64_t>(0xD2511F53u) * c0;
    const uint64_t p1 = static_cast<uint64_t>(0xCD9E8D57u) * c2;
    const uint32_t n0 = static_cast<uint32_t>(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n1 = static_cast<uint32_t>(p1);
    const uint32_t n2 = static_cast<uint32_t>(p0 >> 32) ^ c3 ^ k1;
    const uint32_t n3 = static_cast<uint32_t>(p0);
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return Philox4{c0, c1, c2, c3};
}

__device__ __forceinline__ void box_muller(uint32_t r0, uint32_t r1, float& z0, float& z1) {
  const float u0 = static_cast<float>(r0) * 2.3283064365386963e-10f + 1.1641532182693481e-10f;
  const float u1 = static_cast<float>(r1) * 2.3283064365386963e-10f + 1.1641532182693481e-10f;
  // MUFU paths (lg2 / rsq / sin / cos): |abs err| ~ 4e-7 on a unit-variance draw; the kernel is bound by
  // the Philox integer work, so the transcendental side is kept to one MUFU each
  const float rad = __fsqrt_rn(-1.3862943611198906f * __log2f(u0));
  float s, c;
  __sincosf(6.283185307179586f * u1, &s, &c);
  z0 = rad * c;
  z1 = rad * s;
}

// per-batch parameters read from device memory (CUDA-graph replay: the kernel node is captured once,
// the host rewrites this struct before every replay)
struct NoiseDyn {
  unsigned long long seed;
  unsigned long long first_sample;
  unsigned int stream_id;
  float sigma;
};

struct NoiseParams {
  const float* x;      // [3,S,S]
  const float* eps;    // [B,3,S,S] injected standard draws, or null -> Philox
  uint64_t seed;
  uint32_t stream_id;
  uint64_t first_sample;
  float sigma;
  float mean[4], stdv[4];
  int noise_space, noise_kind, S;
  long long per_img;   // elements per image (C*H*W)
  const NoiseDyn* dyn; // optional override of (seed, first_sample, stream_id, sigma)
};

// noisy, normalised values of 4 consecutive pixels (c, y, x0..x0+3) of sample b (local index)
__device__ __forceinline__ float4 noisy4(const NoiseParams& p, int b, int c, long long e) {
  const float4 xv = __ldg(reinterpret_cast<const float4*>(p.x + e));
  float4 z;
  if (p.eps != nullptr) {
    z = __ldcs(reinterpret_cast<const float4*>(p.eps + static_cast<long long>(b) * p.per_img + e));
  } else {
    const uint64_t sample = p.first_sample + static_cast<uint64_t>(b);
    const Philox4 r = philox4x32_10(static_cast<uint32_t>(e >> 2), static_cast<uint32_t>(sample),
                                    p.stream_id, static_cast<uint32_t>(sample >> 32),
                                    static_cast<uint32_t>(p.seed), static_cast<uint32_t>(p.seed >> 32));
    if (p.noise_kind == CGPT_NOISE_GAUSSIAN) {
      box_muller(r.x, r.y, z.x, z.y);
      box_muller(r.z, r.w, z.z, z.w);
    } else {
      z.x = static_cast<float>(r.x >> 8) * 5.9604644775390625e-08f;
      z.y = static_cast<float>(r.y >> 8) * 5.9604644775390625e-08f;
      z.z = static_cast<float>(r.z >> 8) * 5.9604644775390625e-08f;
      z.w = static_cast<float>(r.w >> 8) * 5.9604644775390625e-08f;
    }
  }
  // smoothing.py:96-97: noise = eps * sigma; batch + noise  (two fp32 roundings, no FMA)
  float4 v;
  v.x = __fadd_rn(xv.x, __fmul_rn(z.x, p.sigma));
  v.y = __fadd_rn(xv.y, __fmul_rn(z.y, p.sigma));
  v.z = __fadd_rn(xv.z, __fmul_rn(z.z, p.sigma));
  v.w = __fadd_rn(xv.w, __fmul_rn(z.w, p.sigma));
  if (p.noise_space == CGPT_SPACE_PIXEL) {
    // transforms.Normalize: (t - mean) / std
    const float m = p.mean[c], s = p.stdv[c];
    v.x = __fdiv_rn(__fsub_rn(v.x, m), s);
    v.y = __fdiv_rn(__fsub_rn(v.y, m), s);
    v.z = __fdiv_rn(__fsub_rn(v.z, m), s);
    v.w = __fdiv_rn(__fsub_rn(v.w, m), s);
  }
  return v;
}

__global__ void __launch_bounds__(256) noise_patchify_kernel(NoiseParams p, __nv_bfloat16* __restrict__ out,
                                                             long long ld_out) {
  if (p.dyn != nullptr) {
    p.seed = p.dyn->seed; p.first_sample = p.dyn->first_sample; p.stream_id = p.dyn->stream_id; p.sigma = p.dyn->sigma;
  }
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __nv_bfloat16* tile = reinterpret_cast<__nv_bfloat16*>(smem_raw);  // [G][592]
  const int S = p.S;
  const int G = S / PATCH;      // patches per row (16 at 224 px)
  const int b = blockIdx.y;
  const int py = blockIdx.x;
  const int groups_per_row = S / 4;
  const int total_groups = 3 * PATCH * groups_per_row;

  // zero the 4 pad columns of every patch row
  for (int i = threadIdx.x; i < G * (PATCH_K_PAD - PATCH_K); i += blockDim.x)
    tile[(i >> 2) * PATCH_K_PAD + PATCH_K + (i & 3)] = __float2bfloat16(0.f);

  for (int g = threadIdx.x; g < total_groups; g += blockDim.x) {
    const int xg = g % groups_per_row;
    const int row = g / groups_per_row;  // 0..41
    const int ky = row % PATCH;
    const int c = row / PATCH;
    const int x0 = xg * 4;
    const float4 v = noisy4(p, b, c, (static_cast<long long>(c) * S + py * PATCH + ky) * S + x0);
    const float vv[4] = {v.x, v.y, v.z, v.w};
    const int col_base = c * (PATCH * PATCH) + ky * PATCH;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int x = x0 + j;
      const int px = x / PATCH;
      const int kx = x - px * PATCH;
      tile[px * PATCH_K_PAD + col_base + kx] = __float2bfloat16(vv[j]);
    }
  }
  __syncthreads();
  // the G patch rows of this patch-row are consecutive output rows
  const long long row0 = (static_cast<long long>(b) * G + py) * G;
  if (ld_out == PATCH_K_PAD) {
    uint4* dst = reinterpret_cast<uint4*>(out + row0 * ld_out);
    const uint4* src = reinterpret_cast<const uint4*>(tile);
    const int n16 = G * PATCH_K_PAD * 2 / 16;
    for (int i = threadIdx.x; i < n16; i += blockDim.x) dst[i] = src[i];
  } else {
    const int per_row = PATCH_K_PAD * 2 / 16;  // 74
    for (int i = threadIdx.x; i < G * per_row; i += blockDim.x) {
      const int r = i / per_row, q = i - r * per_row;
      *(reinterpret_cast<uint4*>(out + (row0 + r) * ld_out) + q) =
          reinterpret_cast<const uint4*>(tile + r * PATCH_K_PAD)[q];
    }
  }
}

// generic path (any torch.nn.Module base classifier): out[b] = x + sigma*eps (+Normalize), fp32 NCHW
__global__ void __launch_bounds__(256) noise_image_kernel(NoiseParams p, long long hw,
                                                          float* __restrict__ out) {
  const long long per_img = p.per_img;
  const long long groups = per_img / 4;
  const int b = blockIdx.y;
  for (long long g = blockIdx.x * blockDim.x + threadIdx.x; g < groups;
       g += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long e = g * 4;
    const int c = static_cast<int>(e / hw);  // hw % 4 == 0: a group never straddles channels
    const float4 v = noisy4(p, b, c, e);
    __stcs(reinterpret_cast<float4*>(out + b * per_img + e), v);
  }
}

static int fill_params(NoiseParams& p, const float* x, const float* eps, uint64_t seed,
                       uint32_t stream_id, uint64_t first_sample, int B, float sigma,
                       const float* mean3, const float* std3, int noise_space, int noise_kind,
                       int img_size, int channels) {
  CGPT_REQUIRE(x != nullptr, "noise: x is null");
  CGPT_REQUIRE(B > 0, "noise: B must be positive (got %d)", B);
  CGPT_REQUIRE(channels >= 1 && channels <= 4, "noise: 1..4 channels supported (got %d)", channels);
  CGPT_REQUIRE(noise_space == CGPT_SPACE_NORMALIZED || noise_space == CGPT_SPACE_PIXEL,
               "noise: bad noise_space %d", noise_space);
  CGPT_REQUIRE(noise_kind == CGPT_NOISE_GAUSSIAN || noise_kind == CGPT_NOISE_UNIFORM,
               "noise: bad noise_kind %d", noise_kind);
  CGPT_REQUIRE(noise_space == CGPT_SPACE_NORMALIZED || (mean3 && std3),
               "noise: PIXEL space needs mean/std");
  p.dyn = nullptr;
  p.x = x; p.eps = eps; p.seed = seed; p.stream_id = stream_id; p.first_sample = first_sample;
  p.sigma = sigma; p.noise_space = noise_space; p.noise_kind = noise_kind; p.S = img_size;
  for (int i = 0; i < 4; ++i) {
    p.mean[i] = (mean3 && i < channels) ? mean3[i] : 0.f;
    p.stdv[i] = (std3 && i < channels) ? std3[i] : 1.f;
  }
  return 0;
}

int noise_patchify(const float* x, const float* eps, uint64_t seed, uint32_t stream_id,
                   uint64_t first_sample, int B, float sigma, const float* mean3,
                   const float* std3, int noise_space, int noise_kind, int img_size, void* out,
                   long long ld_out, const void* dyn, cudaStream_t stream) {
  NoiseParams p;
  if (int rc = fill_params(p, x, eps, seed, stream_id, first_sample, B, sigma, mean3, std3,
                           noise_space, noise_kind, img_size, 3))
    return rc;
  p.dyn = reinterpret_cast<const NoiseDyn*>(dyn);
  CGPT_REQUIRE(img_size > 0 && img_size % PATCH == 0 && img_size % 4 == 0,
               "noise_patchify: image size %d must be a multiple of 14 and of 4", img_size);
  p.per_img = 3LL * img_size * img_size;
  CGPT_REQUIRE(out != nullptr && ld_out >= PATCH_K_PAD && ld_out % 8 == 0,
               "noise_patchify: ld_out must be >= 592 and a multiple of 8 (got %lld)", ld_out);
  CGPT_REQUIRE(B <= 65535, "noise_patchify: at most 65535 samples per launch");
  const int G = img_size / PATCH;
  const size_t smem = static_cast<size_t>(G) * PATCH_K_PAD * 2;
  static bool configured = false;
  if (!configured) {
    CGPT_CHECK_CUDA(cudaFuncSetAttribute(noise_patchify_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    configured = true;
  }
  CGPT_REQUIRE(smem <= 96 * 1024, "noise_patchify: image too large");
  noise_patchify_kernel<<<dim3(G, B), 256, smem, stream>>>(p, reinterpret_cast<__nv_bfloat16*>(out), ld_out);
  CGPT_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int noise_image(const float* x, const float* eps, uint64_t seed, uint32_t stream_id,
                uint64_t first_sample, int B, float sigma, const float* mean3, const float* std3,
                int noise_space, int noise_kind, int channels, int height, int width, float* out,
                cudaStream_t stream) {
  NoiseParams p;
  if (int rc = fill_params(p, x, eps, seed, stream_id, first_sample, B, sigma, mean3, std3,
                           noise_space, noise_kind, width, channels))
    return rc;
  CGPT_REQUIRE(out != nullptr, "noise_image: out is null");
  CGPT_REQUIRE(B <= 65535, "noise_image: at most 65535 samples per launch");
  const long long hw = static_cast<long long>(height) * width;
  CGPT_REQUIRE(hw > 0 && hw % 4 == 0, "noise_image: H*W must be a positive multiple of 4 (got %lld)", hw);
  p.per_img = hw * channels;
  const long long groups = p.per_img / 4;
  int gx = static_cast<int>((groups + 255) / 256);
  if (gx > 1024) gx = 1024;
  noise_image_kernel<<<dim3(gx, B), 256, 0, stream>>>(p, hw, out);
  CGPT_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace cgpt
