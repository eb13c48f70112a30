// K16-K18: answer-token-sequence -> label hash, label histogram, and the statistics tail of
// Smooth.certify / Smooth.predict, all on the device so the Monte-Carlo loop needs no per-batch
// host sync (the reference does predictions.cpu().numpy() + a Python counting loop per batch,
// randomized_smoothing/smoothing.py:98,101-105).
//
//   answer_labels : generated ids -> canonical token sequence -> 64-bit FNV-1a -> open-addressing
//                   table lookup -> class id (unknown answers -> `other_label`).  Canonicalisation
//                   restates minigpt_base.py:438-446 at token level: stop at the first EOS,
//                   drop special ids (<unk>=0, <s>=1, </s>=2: decode(skip_special_tokens=True)).
//   argmax_rows   : logits.argmax(1) (+ top-2 margin) for generic classifiers / the lm_head
//   label_hist    : counts[label] += 1 with warp-aggregated (match.any) 64-bit atomics
//   certify_tail  : cAHat = argmax(counts_sel) (lowest index on ties, smoothing.py:46),
//                   nA = counts_est[cAHat], pABar = BetaInv(alpha; nA, n-nA+1)  (:51,:117),
//                   pABar < 0.5 ? (ABSTAIN, 0) : (cAHat, sigma * PhiInv(pABar))  (:52-56)
//   predict_tail  : top-2 counts, two-sided exact binomial test p=0.5 vs alpha (:73-79)
#include <math.h>
#include "common.cuh"
#include "ops.h"

namespace cgpt {

// ------------------------------------------------------------------ answer hash
__host__ __device__ __forceinline__ uint64_t fnv1a_push(uint64_t h, uint32_t v) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    h ^= (v >> (8 * i)) & 0xffu;
    h *= 0x100000001b3ull;
  }
  return h;
}

__global__ void answer_labels_kernel(const int* __restrict__ ids, int B, int max_new, int ld_ids,
                                     int eos_id, const uint64_t* __restrict__ keys,
                                     const int* __restrict__ vals, int cap_mask, int other_label,
                                     int* __restrict__ labels) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  uint64_t h = 0xcbf29ce484222325ull;
  int len = 0;
  for (int t = 0; t < max_new; ++t) {
    const int id = ids[static_cast<long long>(b) * ld_ids + t];
    if (id == eos_id) break;
    if (id == 0 || id == 1 || id == 2) continue;
    h = fnv1a_push(h, static_cast<uint32_t>(id));
    ++len;
  }
  h = fnv1a_push(h, static_cast<uint32_t>(len));
  if (h == 0) h = 1;
  int label = other_label;
  uint32_t slot = static_cast<uint32_t>(h ^ (h >> 32)) & cap_mask;
  for (int probe = 0; probe <= cap_mask; ++probe) {
    const uint64_t k = keys[slot];
    if (k == h) { label = vals[slot]; break; }
    if (k == 0) break;
    slot = (slot + 1) & cap_mask;
  }
  labels[b] = label;
}

// ------------------------------------------------------------------ row argmax (+ margin)
__global__ void __launch_bounds__(256) argmax_rows_kernel(const float* __restrict__ logits, int rows,
                                                          int cols, long long ld, int suppress_col,
                                                          int* __restrict__ out_idx,
                                                          float* __restrict__ out_margin) {
  const int r = blockIdx.x;
  if (r >= rows) return;
  const float* row = logits + static_cast<long long>(r) * ld;
  float best = -INFINITY, second = -INFINITY;
  int bi = 0x7fffffff;
  // (v, c) beats (b, i): larger value, lowest index on ties; NaN counts as the maximum (torch.argmax), so a
  // poisoned row still yields an index inside [0, cols) - the next decode step gathers an embedding row with it
  auto beats = [](float v, int c, float b, int i) {
    const bool vn = v != v, bn = b != b;
    if (vn || bn) return vn && (!bn || c < i);
    return v > b || (v == b && c < i);
  };
  for (int c = threadIdx.x; c < cols; c += blockDim.x) {
    float v = row[c];
    if (c == suppress_col) v = -INFINITY;
    if (beats(v, c, best, bi)) { second = best; best = v; bi = c; }
    else if (v > second) second = v;
  }
  // warp then block reduction of (best, idx, second); lowest index wins ties (torch.argmax)
  auto merge = [&beats](float& b1, int& i1, float& s1, float b2, int i2, float s2) {
    if (beats(b2, i2, b1, i1)) { s1 = fmaxf(b1, s2); b1 = b2; i1 = i2; }
    else { s1 = fmaxf(s1, b2); }
  };
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float b2 = __shfl_xor_sync(0xffffffffu, best, o);
    const int i2 = __shfl_xor_sync(0xffffffffu, bi, o);
    const float s2 = __shfl_xor_sync(0xffffffffu, second, o);
    merge(best, bi, second, b2, i2, s2);
  }
  __shared__ float sb[8], ss[8];
  __shared__ int si[8];
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { sb[w] = best; si[w] = bi; ss[w] = second; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < (blockDim.x >> 5); ++k) merge(best, bi, second, sb[k], si[k], ss[k]);
    out_idx[r] = (bi >= 0 && bi < cols) ? bi : 0;
    if (out_margin) out_margin[r] = best - second;
  }
}

// ------------------------------------------------------------------ greedy bookkeeping
// HF greedy search row update: finished rows emit pad; a row finishes when it emits EOS.
__global__ void greedy_step_kernel(const int* __restrict__ next_idx, int B, int* __restrict__ finished,
                                   int* __restrict__ ids_out, int ld, int t, int eos_id, int pad_id,
                                   int* __restrict__ unfinished_count) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int fin = finished[b];
  const int tok = fin ? pad_id : next_idx[b];
  ids_out[static_cast<long long>(b) * ld + t] = tok;
  const int now = fin | (tok == eos_id);
  finished[b] = now;
  if (!now && unfinished_count) atomicAdd(unfinished_count, 1);
}

// ------------------------------------------------------------------ histogram
__global__ void __launch_bounds__(256) label_hist_kernel(const int* __restrict__ labels, int B,
                                                         int num_classes,
                                                         unsigned long long* __restrict__ counts,
                                                         int* __restrict__ invalid) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = i < B;
  int label = active ? labels[i] : -1;
  const bool ok = active && label >= 0 && label < num_classes;
  if (active && !ok && invalid) atomicAdd(invalid, 1);
  const unsigned mask = __ballot_sync(0xffffffffu, ok);
  if (ok) {
    const unsigned peers = __match_any_sync(mask, label);
    const int leader = __ffs(peers) - 1;
    if ((threadIdx.x & 31) == leader)
      atomicAdd(&counts[label], static_cast<unsigned long long>(__popc(peers)));
  }
}

// ------------------------------------------------------------------ fp64 statistics
// Phi^-1: Wichura's AS241 PPND16 (relative accuracy ~1e-16)
__device__ double ppnd16(double p) {
  const double q = p - 0.5;
  double r, val;
  if (fabs(q) <= 0.425) {
    r = 0.180625 - q * q;
    val = q * (((((((2.5090809287301226727e3 * r + 3.3430575583588128105e4) * r + 6.7265770927008700853e4) * r +
                   4.5921953931549871457e4) * r + 1.3731693765509461125e4) * r + 1.9715909503065514427e3) * r +
                 1.3314166789178437745e2) * r + 3.3871328727963666080e0) /
          (((((((5.2264952788528545610e3 * r + 2.8729085735721942674e4) * r + 3.9307895800092710610e4) * r +
               2.1213794301586595867e4) * r + 5.3941960214247511077e3) * r + 6.8718700749205790830e2) * r +
            4.2313330701600911252e1) * r + 1.0);
    return val;
  }
  r = q < 0 ? p : 1.0 - p;
  r = sqrt(-log(r));
  if (r <= 5.0) {
    r -= 1.6;
    val = (((((((7.74545014278341407640e-4 * r + 2.27238449892691845833e-2) * r + 2.41780725177450611770e-1) * r +
               1.27045825245236838258e0) * r + 3.64784832476320460504e0) * r + 5.76949722146069140550e0) * r +
            4.63033784615654529590e0) * r + 1.42343711074968357734e0) /
          (((((((1.05075007164441684324e-9 * r + 5.47593808499534494600e-4) * r + 1.51986665636164571966e-2) * r +
               1.48103976427480074590e-1) * r + 6.89767334985100004550e-1) * r + 1.67638483018380384940e0) * r +
            2.05319162663775882187e0) * r + 1.0);
  } else {
    r -= 5.0;
    val = (((((((2.01033439929228813265e-7 * r + 2.71155556874348757815e-5) * r + 1.24266094738807843860e-3) * r +
               2.65321895265761230930e-2) * r + 2.96560571828504891230e-1) * r + 1.78482653991729133580e0) * r +
            5.46378491116411436990e0) * r + 6.65790464350110377720e0) /
          (((((((2.04426310338993978564e-15 * r + 1.42151175831644588870e-7) * r + 1.84631831751005468180e-5) * r +
               7.86869131145613259100e-4) * r + 1.48753612908506148525e-2) * r + 1.36929880922735805310e-1) * r +
            5.99832206555887937690e-1) * r + 1.0);
  }
  return q < 0 ? -val : val;
}

// block-wide deterministic sum (fixed tree), result valid in every thread
__device__ double block_sum(double v, double* sh) {
  __syncthreads();
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  for (int k = 0; k < (blockDim.x >> 5); ++k) t += sh[k];
  return t;
}

// P[Bin(n, p) >= k0]  =  I_p(k0, n-k0+1), summed term by term in fp64
__device__ double binom_upper_tail(long long n, long long k0, double p, double* sh) {
  const double lp = log(p), lq = log1p(-p);
  const double lgn = lgamma(static_cast<double>(n) + 1.0);
  double acc = 0.0;
  for (long long k = k0 + threadIdx.x; k <= n; k += blockDim.x) {
    const double lt = lgn - lgamma(static_cast<double>(k) + 1.0) - lgamma(static_cast<double>(n - k) + 1.0) +
                      static_cast<double>(k) * lp + static_cast<double>(n - k) * lq;
    acc += exp(lt);
  }
  return block_sum(acc, sh);
}

// Clopper-Pearson lower bound: the p with P[Bin(n,p) >= nA] = alpha
__device__ double clopper_pearson_lower(long long nA, long long n, double alpha, double* sh) {
  if (nA <= 0) return 0.0;
  if (nA >= n) return pow(alpha, 1.0 / static_cast<double>(n));
  double lo = 0.0, hi = 1.0;
  for (int it = 0; it < 200; ++it) {
    const double mid = 0.5 * (lo + hi);
    if (mid <= lo || mid >= hi) break;
    const double f = binom_upper_tail(n, nA, mid, sh);
    if (f < alpha) lo = mid; else hi = mid;
  }
  return 0.5 * (lo + hi);
}

__global__ void __launch_bounds__(256) certify_tail_kernel(const long long* __restrict__ counts_sel,
                                                           const long long* __restrict__ counts_est,
                                                           int num_classes, long long n, double alpha,
                                                           double sigma, const double* __restrict__ lut,
                                                           long long img_stride, int* __restrict__ out_label,
                                                           double* __restrict__ out_stats) {
  __shared__ double sh[8];
  __shared__ long long s_best[256];
  __shared__ int s_idx[256];
  // one block per image: image k's count vectors are `img_stride` elements further on, its outputs 3 further on
  counts_sel += blockIdx.x * img_stride;
  counts_est += blockIdx.x * img_stride;
  out_label += blockIdx.x * 3;
  out_stats += blockIdx.x * 3;
  long long best = -1;
  int bi = 0x7fffffff;
  for (int c = threadIdx.x; c < num_classes; c += blockDim.x) {
    const long long v = counts_sel[c];
    if (v > best) { best = v; bi = c; }  // ascending c per thread: first max kept
  }
  s_best[threadIdx.x] = best; s_idx[threadIdx.x] = bi;
  __syncthreads();
  for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      const long long b2 = s_best[threadIdx.x + o];
      const int i2 = s_idx[threadIdx.x + o];
      if (b2 > s_best[threadIdx.x] || (b2 == s_best[threadIdx.x] && i2 < s_idx[threadIdx.x])) {
        s_best[threadIdx.x] = b2; s_idx[threadIdx.x] = i2;
      }
    }
    __syncthreads();
  }
  const int cA = s_idx[0];
  const long long nA = counts_est[cA];
  // lut (optional): [pABar(nA) | Phi^-1(pABar(nA))] for nA = 0..n, tabulated on the host once per (n, alpha) with the
  // SciPy calls the reference makes per image (smoothing.py:55,117) -> pABar and radius bit-identical to the
  // reference (the single fp64 multiply by sigma is IEEE-exact on both sides).  Without it: device bisection + AS241.
  const bool use_lut = lut != nullptr && nA >= 0 && nA <= n;
  const double pABar = use_lut ? lut[nA] : clopper_pearson_lower(nA, n, alpha, sh);
  if (threadIdx.x == 0) {
    if (pABar < 0.5) {
      out_label[0] = -1;
      out_stats[0] = 0.0;
    } else {
      out_label[0] = cA;
      out_stats[0] = __dmul_rn(sigma, use_lut ? lut[n + 1 + nA] : ppnd16(pABar));
    }
    out_label[1] = cA;
    out_stats[1] = pABar;
    out_stats[2] = static_cast<double>(nA);
  }
}

__global__ void __launch_bounds__(256) predict_tail_kernel(const long long* __restrict__ counts,
                                                           int num_classes, double alpha,
                                                           int* __restrict__ out_label,
                                                           double* __restrict__ out_stats) {
  __shared__ double sh[8];
  __shared__ long long s_best[256];
  __shared__ int s_idx[256];
  int top[2] = {-1, -1};
  long long topc[2] = {0, 0};
  for (int pass = 0; pass < 2; ++pass) {
    long long best = -1;
    int bi = 0x7fffffff;
    for (int c = threadIdx.x; c < num_classes; c += blockDim.x) {
      if (pass == 1 && c == top[0]) continue;
      const long long v = counts[c];
      if (v > best) { best = v; bi = c; }
    }
    s_best[threadIdx.x] = best; s_idx[threadIdx.x] = bi;
    __syncthreads();
    for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
      if (threadIdx.x < o) {
        const long long b2 = s_best[threadIdx.x + o];
        const int i2 = s_idx[threadIdx.x + o];
        if (b2 > s_best[threadIdx.x] || (b2 == s_best[threadIdx.x] && i2 < s_idx[threadIdx.x])) {
          s_best[threadIdx.x] = b2; s_idx[threadIdx.x] = i2;
        }
      }
      __syncthreads();
    }
    top[pass] = s_idx[0];
    topc[pass] = s_best[0] < 0 ? 0 : s_best[0];
    __syncthreads();
  }
  // two-sided exact binomial test with p = 0.5 (symmetric): 2 * P[X >= max(c1,c2)], capped at 1
  const long long c1 = topc[0], c2 = topc[1], m = c1 + c2;
  double pval = 1.0;
  if (c1 != c2 && m > 0) {
    const long long hi = c1 > c2 ? c1 : c2;
    const double tail = binom_upper_tail(m, hi, 0.5, sh);
    pval = fmin(1.0, 2.0 * tail);
  }
  if (threadIdx.x == 0) {
    out_label[0] = (pval > alpha) ? -1 : top[0];
    out_label[1] = top[0];
    out_label[2] = top[1];
    out_stats[0] = pval;
    out_stats[1] = static_cast<double>(c1);
    out_stats[2] = static_cast<double>(c2);
  }
}

// ------------------------------------------------------------------ host wrappers
int answer_labels(const int* ids, int B, int max_new, int ld_ids, int eos_id, const uint64_t* keys,
                  const int* vals, int capacity, int other_label, int* labels, cudaStream_t stream) {
  CGPT_REQUIRE(B > 0 && max_new > 0 && ld_ids >= max_new, "answer_labels: bad shape B=%d max_new=%d ld=%d",
               B, max_new, ld_ids);
  CGPT_REQUIRE(capacity > 0 && (capacity & (capacity - 1)) == 0,
               "answer_labels: table capacity must be a power of two (got %d)", capacity);
  answer_labels_kernel<<<(B + 127) / 128, 128, 0, stream>>>(ids, B, max_new, ld_ids, eos_id, keys, vals,
                                                           capacity - 1, other_label, labels);
  CGPT_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int argmax_rows(const float* logits, int rows, int cols, long long ld, int suppress_col, int* out_idx,
                float* out_margin, cudaStream_t stream) {
  CGPT_REQUIRE(rows > 0 && cols > 0 && ld >= cols, "argmax_rows: bad shape rows=%d cols=%d", rows, cols);
  argmax_rows_kernel<<<rows, 256, 0, stream>>>(logits, rows, cols, ld, suppress_col, out_idx, out_margin);
  CGPT_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int greedy_step(const int* next_idx, int B, int* finished, int* ids_out, int ld, int t, int eos_id,
                int pad_id, int* unfinished_count, cudaStream_t stream) {
  CGPT_REQUIRE(B > 0 && t >= 0 && t < ld, "greedy_step: bad arguments B=%d t=%d ld=%d", B, t, ld);
  greedy_step_kernel<<<(B + 127) / 128, 128, 0, stream>>>(next_idx, B, finished, ids_out, ld, t, eos_id,
                                                         pad_id, unfinished_count);
  CGPT_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int label_hist(const int* labels, int B, int num_classes, long long* counts, int* invalid,
               cudaStream_t stream) {
  CGPT_REQUIRE(B >= 0 && num_classes > 0, "label_hist: bad shape B=%d classes=%d", B, num_classes);
  if (B == 0) return 0;
  label_hist_kernel<<<(B + 255) / 256, 256, 0, stream>>>(
      labels, B, num_classes, reinterpret_cast<unsigned long long*>(counts), invalid);
  CGPT_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int certify_tail(const long long* counts_sel, const long long* counts_est, int num_classes, long long n,
                 double alpha, double sigma, const double* lut, int* out_label, double* out_stats,
                 cudaStream_t stream) {
  CGPT_REQUIRE(num_classes > 0 && n > 0 && alpha > 0.0 && alpha < 1.0,
               "certify_tail: bad arguments classes=%d n=%lld alpha=%g", num_classes, n, alpha);
  certify_tail_kernel<<<1, 256, 0, stream>>>(counts_sel, counts_est, num_classes, n, alpha, sigma, lut, 0,
                                             out_label, out_stats);
  CGPT_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

// K images at once: image k's selection / estimation vectors at counts + k * 2 * num_classes (+ num_classes),
// out_label[3k..] / out_stats[3k..] as in certify_tail
int certify_tail_batch(const long long* counts, int images, int num_classes, long long n, double alpha, double sigma,
                       const double* lut, int* out_label, double* out_stats, cudaStream_t stream) {
  CGPT_REQUIRE(images > 0 && num_classes > 0 && n > 0 && alpha > 0.0 && alpha < 1.0,
               "certify_tail_batch: bad arguments images=%d classes=%d n=%lld alpha=%g", images, num_classes, n, alpha);
  certify_tail_kernel<<<images, 256, 0, stream>>>(counts, counts + num_classes, num_classes, n, alpha, sigma, lut,
                                                  2LL * num_classes, out_label, out_stats);
  CGPT_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

// labels of a multi-image pass: row r belongs to image r / per_image and is global sample first + r % per_image of it;
// samples below `boundary` are counted into the image's selection vector, the others into its estimation vector
__global__ void __launch_bounds__(256) label_hist_images_kernel(const int* __restrict__ labels, int rows, int per_image,
                                                                long long first, long long boundary, int num_classes,
                                                                unsigned long long* __restrict__ counts,
                                                                int* __restrict__ invalid) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  const int label = labels[i];
  if (label < 0 || label >= num_classes) {
    if (invalid) atomicAdd(invalid, 1);
    return;
  }
  const int img = i / per_image;
  const long long sample = first + (i - img * per_image);
  const int vec = (boundary >= 0 && sample >= boundary) ? 1 : 0;
  atomicAdd(&counts[(static_cast<long long>(img) * 2 + vec) * num_classes + label], 1ull);
}

int label_hist_images(const int* labels, int rows, int per_image, long long first, long long boundary, int num_classes,
                      long long* counts, int* invalid, cudaStream_t stream) {
  CGPT_REQUIRE(rows > 0 && per_image > 0 && rows % per_image == 0 && num_classes > 0,
               "label_hist_images: bad shape rows=%d per_image=%d", rows, per_image);
  label_hist_images_kernel<<<(rows + 255) / 256, 256, 0, stream>>>(labels, rows, per_image, first, boundary, num_classes,
                                                                  reinterpret_cast<unsigned long long*>(counts), invalid);
  CGPT_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int predict_tail(const long long* counts, int num_classes, double alpha, int* out_label,
                 double* out_stats, cudaStream_t stream) {
  CGPT_REQUIRE(num_classes > 0 && alpha > 0.0 && alpha < 1.0, "predict_tail: bad arguments");
  predict_tail_kernel<<<1, 256, 0, stream>>>(counts, num_classes, alpha, out_label, out_stats);
  CGPT_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

uint64_t answer_hash_host(const int* ids, int n, int eos_id) {
  uint64_t h = 0xcbf29ce484222325ull;
  int len = 0;
  for (int t = 0; t < n; ++t) {
    const int id = ids[t];
    if (id == eos_id) break;
    if (id == 0 || id == 1 || id == 2) continue;
    h = fnv1a_push(h, static_cast<uint32_t>(id));
    ++len;
  }
  h = fnv1a_push(h, static_cast<uint32_t>(len));
  if (h == 0) h = 1;
  return h;
}


// ------------------------------------------------------------------ cosine of feature rows against one target
__global__ void __launch_bounds__(128) cosine_rows_kernel(const float* __restrict__ feats, long long ld, int rows, int D,
                                                          const float* __restrict__ target, float* __restrict__ scores) {
  const int r = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  const float* f = feats + r * ld;
  float dot = 0.f, nf = 0.f, nt = 0.f;
  for (int c = lane; c < D; c += 32) {
    const float a = f[c], b = __ldg(target + c);
    dot = fmaf(a, b, dot);
    nf = fmaf(a, a, nf);
    nt = fmaf(b, b, nt);
  }
  dot = warp_sum(dot); nf = warp_sum(nf); nt = warp_sum(nt);
  if (lane == 0) scores[r] = dot / fmaxf(sqrtf(nf) * sqrtf(nt), 1e-12f);
}

int cosine_rows(const float* feats, long long ld, int rows, int D, const float* target, float* scores,
                cudaStream_t stream) {
  CGPT_REQUIRE(feats && target && scores && rows > 0 && D > 0, "cosine_rows: bad arguments rows=%d D=%d", rows, D);
  cosine_rows_kernel<<<(rows + 3) / 4, 128, 0, stream>>>(feats, ld, rows, D, target, scores);
  CGPT_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}


// ------------------------------------------------------------------ language-model loss (modeling_llama.py:101-123)
// loss[r] = logsumexp(z) - (1 - eps) * z[target[r]] - eps * mean(z)   (z = logits[r, :], fp32; eps = label smoothing,
// CrossEntropyLoss(label_smoothing=0.1) at modeling_llama.py:107; eps = 0 is the plain loss); 0 and not counted when
// target[r] < 0 = ignore_index
__global__ void __launch_bounds__(256) ce_rows_kernel(const float* __restrict__ logits, long long ld, int cols,
                                                      const int* __restrict__ targets, float* __restrict__ loss, float eps) {
  __shared__ float red[8];
  __shared__ float red2[8];
  const int r = blockIdx.x;
  const int t = targets[r];
  if (t < 0 || t >= cols) {          // uniform per block
    if (threadIdx.x == 0) loss[r] = 0.f;
    return;
  }
  const float* row = logits + r * ld;
  float mx = -INFINITY, zs = 0.f;
  for (int c = threadIdx.x; c < cols; c += 256) { mx = fmaxf(mx, row[c]); zs += row[c]; }
  mx = warp_max(mx);
  zs = warp_sum(zs);
  if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = mx; red2[threadIdx.x >> 5] = zs; }
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
  if (threadIdx.x == 0) {
    float tot = 0.f, ztot = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { tot += red[i]; ztot += red2[i]; }   // fixed order
    loss[r] = logf(tot) + mx - (1.f - eps) * row[t] - eps * (ztot / static_cast<float>(cols));
  }
}
// out[0] = mean of loss[r] over targets[r] >= 0 (CrossEntropyLoss reduction='mean'), out[1] = their count.
// One block, fixed summation order: the result does not depend on scheduling.
__global__ void __launch_bounds__(256) masked_mean_kernel(const float* __restrict__ loss, const int* __restrict__ targets,
                                                          int rows, int cols, float* __restrict__ out) {
  __shared__ double s_sum[256];
  __shared__ int s_cnt[256];
  double acc = 0.0;
  int cnt = 0;
  for (int r = threadIdx.x; r < rows; r += 256) {
    const int t = targets[r];
    if (t >= 0 && t < cols) { acc += static_cast<double>(loss[r]); ++cnt; }
  }
  s_sum[threadIdx.x] = acc;
  s_cnt[threadIdx.x] = cnt;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) { s_sum[threadIdx.x] += s_sum[threadIdx.x + o]; s_cnt[threadIdx.x] += s_cnt[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    out[0] = s_cnt[0] > 0 ? static_cast<float>(s_sum[0] / s_cnt[0]) : 0.f;
    out[1] = static_cast<float>(s_cnt[0]);
  }
}

int ce_loss(const float* logits, long long ld, int rows, int cols, const int* targets, float* token_loss,
            float* mean_count, float label_smoothing, cudaStream_t stream) {
  CGPT_REQUIRE(logits && targets && token_loss && rows > 0 && cols > 0, "ce_loss: bad arguments rows=%d cols=%d", rows, cols);
  CGPT_REQUIRE(label_smoothing >= 0.f && label_smoothing < 1.f, "ce_loss: label_smoothing %f outside [0, 1)", label_smoothing);
  ce_rows_kernel<<<rows, 256, 0, stream>>>(logits, ld, cols, targets, token_loss, label_smoothing);
  CGPT_CHECK_CUDA(cudaGetLastError());
  count_launch();
  if (mean_count != nullptr) {
    masked_mean_kernel<<<1, 256, 0, stream>>>(token_loss, targets, rows, cols, mean_count);
    CGPT_CHECK_CUDA(cudaGetLastError());
    count_launch();
  }
  return 0;
}

}  // namespace cgpt
