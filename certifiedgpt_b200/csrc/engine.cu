// Native engine: host-side C++ orchestration of the sm_100a kernels behind one handle
// (include/cgpt.h, "native engine").  One call of cgpt_noisy_labels is one batch of the reference's
// hot loop (smoothing.py:95-97 -> MiniGPT4.encode_img minigpt4.py:121-149 -> MiniGPTBase.generate
// minigpt_base.py:374-448 -> answer label); cgpt_sample_noise / cgpt_certify / cgpt_predict are
// Smooth._sample_noise / certify / predict (smoothing.py:29-117) with no Python in the loop.
// The kernel sequence of a batch is captured once per batch size into CUDA graphs and replayed; the
// per-batch noise parameters live in a 24-byte device struct rewritten before each replay.
// Device memory is caller-owned: weights are bound by pointer, all activations / KV cache live in one
// workspace the caller allocates (cgpt_workspace_bytes).
#include <dlfcn.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <new>

#include <string>
#include <unordered_map>
#include <vector>

#include "common.cuh"
#include "ops.h"

namespace cgpt {
namespace {

#define CGPT_TRY(expr)              \
  do {                              \
    if (int _rc = (expr)) return _rc; \
  } while (0)

// ---------------------------------------------------------------- small utility kernels
__global__ void fill_u32_kernel(uint32_t* __restrict__ p, long long n, uint32_t v) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i < n) p[i] = v;
}
// dst[r * dst_ld] = src[r * src_ld]  (4-byte elements): margin[:, t] = mcol ; cur = ids[:, t-1]
__global__ void copy_col_u32_kernel(uint32_t* __restrict__ dst, long long dst_ld, const uint32_t* __restrict__ src,
                                    long long src_ld, int rows) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < rows) dst[r * dst_ld] = src[r * src_ld];
}
// out[b * row_stride + c] = vec[c]: the cls token (+ pos_embed[0]) row of every sample (eva_vit.py:337-340)
__global__ void set_rows_f32_kernel(float* __restrict__ out, long long row_stride, const float* __restrict__ vec,
                                    int D) {
  float* o = out + blockIdx.x * row_stride;
  for (int c = threadIdx.x; c < D; c += blockDim.x) o[c] = vec[c];
}

// dst[b, :] = src[b * T + T - 1, :]  (rows of `row_bytes` bytes, multiple of 16): the last prompt position of
// every sample, the only row the last decoder layer still needs after its attention
__global__ void gather_last_rows_kernel(uint4* __restrict__ dst, const uint4* __restrict__ src, int T, int vec_per_row) {
  const uint4* s = src + (static_cast<long long>(blockIdx.x) * T + T - 1) * vec_per_row;
  uint4* d = dst + static_cast<long long>(blockIdx.x) * vec_per_row;
  for (int c = threadIdx.x; c < vec_per_row; c += blockDim.x) d[c] = s[c];
}
int gather_last_rows(void* dst, const void* src, int T, int B, long long row_bytes, cudaStream_t s) {
  gather_last_rows_kernel<<<B, 256, 0, s>>>(static_cast<uint4*>(dst), static_cast<const uint4*>(src), T,
                                            static_cast<int>(row_bytes / 16));
  CGPT_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

int fill_u32(void* p, long long n, uint32_t v, cudaStream_t s) {
  if (n <= 0) return 0;
  fill_u32_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(static_cast<uint32_t*>(p), n, v);
  CGPT_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}
int copy_col_u32(void* dst, long long dst_ld, const void* src, long long src_ld, int rows, cudaStream_t s) {
  copy_col_u32_kernel<<<(rows + 255) / 256, 256, 0, s>>>(static_cast<uint32_t*>(dst), dst_ld,
                                                         static_cast<const uint32_t*>(src), src_ld, rows);
  CGPT_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}
int set_rows_f32(float* out, long long row_stride, const float* vec, int D, int rows, cudaStream_t s) {
  set_rows_f32_kernel<<<rows, 256, 0, s>>>(out, row_stride, vec, D);
  CGPT_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

// ---------------------------------------------------------------- engine state
struct Tensor {
  const void* p = nullptr;
  long long rows = 0, cols = 0;
  int dtype = 0;
};

struct VitLayer {
  const float *ln1w, *ln1b, *qkvb, *projb, *ln2w, *ln2b, *fc1b, *fc2b;
  const void *qkvw, *projw, *fc1w, *fc2w;
};
struct QfLayer {
  bool cross;
  const void *qkvw, *aow, *cqw, *cow, *fiw, *fow;
  const float *qkvb, *aob, *alnw, *alnb, *cqb, *cob, *clnw, *clnb, *fib, *fob, *flnw, *flnb;
};
struct LlmLayer {
  const void *qkvw, *ow, *guw, *downw;
  const float *n1, *n2;
};

struct Buffers {   // all inside the caller's workspace
  // vision side
  void *patches, *v_res, *v_xn, *v_qkv, *v_att, *v_h, *v_out;
  void *q_h, *q_tmp, *q_qkv, *q_ctx, *q_cq, *q_ckv, *q_inter, *enc_out;
  // language side
  void *l_res, *l_xn, *l_qkv, *l_att, *l_act, *kc, *vc, *kp, *vp, *l_last, *l_logits, *l_res_last, *l_att_last;
  int32_t *ids, *finished, *unfinished, *next, *cur, *labels;
  float *mcol, *margin;
  // per-call
  float* x_static;
  void* dyn;            // NoiseDyn {u64 seed; u64 first_sample; u32 stream_id; f32 sigma}
  long long* counts;    // [2 * num_classes]
  int32_t* invalid;
  int32_t* tail_label;  // [3]
  double* tail_stats;   // [3]
};

struct GraphSet {
  int B, K, space, kind;   // K = images per pass (B / K draws of each)
  int ns;                  // question length and ids buffer the graphs were captured with (cgpt_set_question)
  const void* suffix;
  float mean[3], std[3];
  cudaGraphExec_t g0 = nullptr;
  std::vector<cudaGraphExec_t> steps;
  std::vector<int> nodes;   // libcgpt kernel nodes per graph
};

struct NoiseDyn {
  uint64_t seed;
  uint64_t first_sample;
  uint32_t stream_id;
  float sigma;
};

}  // namespace
}  // namespace cgpt

struct cgpt_engine {
  cgpt_model_config c;
  std::unordered_map<std::string, cgpt::Tensor> w;
  bool resolved = false;
  // derived sizes
  int T = 0, Pn = 0, G = 0, vhd = 0, qhd = 0, lhd = 0, n_cross = 0;
  int P = 0, ns = 0, Tp = 0, cache_rows = 0;   // ns / Tp: the CURRENT question (cgpt_set_question), <= the config's maximum
  int Tp_max = 0;                              // qf_queries + cfg.n_suffix: what the workspace and the KV cache are sized for
  // resolved weights
  const void *patch_w = nullptr, *qf_q0 = nullptr, *ckv_w = nullptr, *proj_w = nullptr, *emb = nullptr,
             *head_w = nullptr;
  const float *patch_b = nullptr, *pos = nullptr, *cls_pos = nullptr, *lnv_w = nullptr, *lnv_b = nullptr,
              *ckv_b = nullptr, *proj_b = nullptr, *llm_norm = nullptr, *rope_cos = nullptr, *rope_sin = nullptr;
  std::vector<cgpt::VitLayer> vit;
  std::vector<cgpt::QfLayer> qf;
  std::vector<cgpt::LlmLayer> llm;
  const int32_t *prefix_ids = nullptr, *suffix_ids = nullptr;
  bool prompt_set = false;
  const uint64_t* table_keys = nullptr;
  const int32_t* table_vals = nullptr;
  int table_cap = 0;
  // workspace
  char* ws = nullptr;
  long long ws_bytes = 0;
  int ws_B = 0;
  bool ws_encoder_only = false;
  cgpt::Buffers b{};
  // graphs
  std::vector<cgpt::GraphSet> graphs;
  cudaStream_t cap_stream = nullptr;
  // pinned host scratch
  int32_t* h_i32 = nullptr;   // [8]
  double* h_f64 = nullptr;    // [8]
  int last_steps = 0;
  int max_images = 1;            // images per pass the workspace serves (cgpt_set_option "max_images", before binding)
  const double* lut = nullptr;   // cgpt_set_radius_lut: [pABar | Phi^-1(pABar)] for (lut_n, lut_alpha), caller-owned
  long long lut_n = 0;
  double lut_alpha = 0.0;
  float label_smoothing = 0.f;   // cgpt_lm_loss: CrossEntropyLoss(label_smoothing) of modeling_llama.py:107 (set_option)
};

namespace cgpt {
namespace {

using Engine = cgpt_engine;

inline long long align_up(long long v, long long a = 256) { return (v + a - 1) / a * a; }

// GEMM wrapper with the field order of _lib.gemm (certifiedgpt_b200/_lib.py)
struct Epi {
  cgpt_gemm_epilogue e;
  Epi(void* out, long long ldo, int odt) {
    memset(&e, 0, sizeof(e));
    e.out = out; e.ldo = ldo; e.out_dtype = odt;
  }
  Epi& bias(const float* b) { e.bias = b; return *this; }
  Epi& resid(const void* r, long long ldr, int dt) { e.resid = r; e.ldr = ldr; e.resid_dtype = dt; return *this; }
  Epi& act(int a) { e.act = a; return *this; }
  Epi& row_add(const float* p, long long ld, int offset) { e.row_add = p; e.ld_row_add = ld; e.row_add_offset = offset; return *this; }
  Epi& rope(const cgpt_gemm_rope* r) { e.rope = r; return *this; }
  Epi& remap(int period, int stride, int offset) { e.row_period = period; e.remap_stride = stride; e.remap_offset = offset; return *this; }
  Epi& headmajor(int T, int heads, int hd) { e.hm_T = T; e.hm_heads = heads; e.hm_hd = hd; return *this; }
};
inline int gemm(const void* A, long long lda, const void* W, int M, int N, int K, const Epi& epi, cudaStream_t s) {
  return gemm_bf16(A, lda, W, K, M, N, K, &epi.e, 0, s);
}
inline __nv_bfloat16* bf(void* p) { return static_cast<__nv_bfloat16*>(p); }

int attn(const void* q, long long ldq, int q_rows, const void* k, const void* v, long long ldkv, int kv_rows,
         void* o, long long ldo, int B, int H, int Tq, int Tk, int hd, float scale, int causal, int decode,
         cudaStream_t s) {
  cgpt_attn_args a;
  memset(&a, 0, sizeof(a));
  a.q = q; a.ldq = ldq; a.q_rows_per_batch = q_rows;
  a.k = k; a.v = v; a.ldk = ldkv; a.ldv = ldkv; a.kv_rows_per_batch = kv_rows;
  a.o = o; a.ldo = ldo;
  a.B = B; a.H = H; a.Tq = Tq; a.Tk = Tk; a.head_dim = hd;
  a.scale = scale; a.causal = causal; a.decode_kernel = decode;
  return attention(&a, s);
}

// ---------------------------------------------------------------- weights
int find(Engine* E, const std::string& name, int dtype, long long rows, long long cols, const void** out) {
  auto it = E->w.find(name);
  CGPT_REQUIRE(it != E->w.end(), "engine: weight '%s' is not bound", name.c_str());
  const Tensor& t = it->second;
  CGPT_REQUIRE(t.dtype == dtype, "engine: weight '%s' has dtype %d, expected %d", name.c_str(), t.dtype, dtype);
  CGPT_REQUIRE(t.rows == rows && t.cols == cols, "engine: weight '%s' is [%lld, %lld], expected [%lld, %lld]",
               name.c_str(), t.rows, t.cols, rows, cols);
  *out = t.p;
  return 0;
}
int mat(Engine* E, const std::string& n, long long rows, long long cols, const void** out) {
  return find(E, n, CGPT_DT_BF16, rows, cols, out);
}
int vec(Engine* E, const std::string& n, long long cols, const float** out) {
  return find(E, n, CGPT_DT_F32, 1, cols, reinterpret_cast<const void**>(out));
}

int resolve(Engine* E) {
  if (E->resolved) return 0;
  const cgpt_model_config& c = E->c;
  const int D = c.vit_dim, Hq = c.qf_hidden, Hl = c.llm_hidden;
  CGPT_TRY(mat(E, "patch.w", D, 592, &E->patch_w));
  CGPT_TRY(vec(E, "patch.b", D, &E->patch_b));
  CGPT_TRY(find(E, "pos", CGPT_DT_F32, E->T, D, reinterpret_cast<const void**>(&E->pos)));
  CGPT_TRY(vec(E, "cls_pos", D, &E->cls_pos));
  E->vit.resize(c.vit_depth);
  for (int i = 0; i < c.vit_depth; ++i) {
    const std::string o = "vit." + std::to_string(i) + ".";
    VitLayer& L = E->vit[i];
    CGPT_TRY(vec(E, o + "ln1.w", D, &L.ln1w));
    CGPT_TRY(vec(E, o + "ln1.b", D, &L.ln1b));
    CGPT_TRY(mat(E, o + "qkv.w", 3 * D, D, &L.qkvw));
    CGPT_TRY(vec(E, o + "qkv.b", 3 * D, &L.qkvb));
    CGPT_TRY(mat(E, o + "proj.w", D, D, &L.projw));
    CGPT_TRY(vec(E, o + "proj.b", D, &L.projb));
    CGPT_TRY(vec(E, o + "ln2.w", D, &L.ln2w));
    CGPT_TRY(vec(E, o + "ln2.b", D, &L.ln2b));
    CGPT_TRY(mat(E, o + "fc1.w", c.vit_mlp, D, &L.fc1w));
    CGPT_TRY(vec(E, o + "fc1.b", c.vit_mlp, &L.fc1b));
    CGPT_TRY(mat(E, o + "fc2.w", D, c.vit_mlp, &L.fc2w));
    CGPT_TRY(vec(E, o + "fc2.b", D, &L.fc2b));
  }
  CGPT_TRY(vec(E, "lnv.w", D, &E->lnv_w));
  CGPT_TRY(vec(E, "lnv.b", D, &E->lnv_b));
  CGPT_TRY(mat(E, "qf.q0", c.qf_queries, Hq, &E->qf_q0));
  E->qf.resize(c.qf_layers);
  for (int i = 0; i < c.qf_layers; ++i) {
    const std::string o = "qf." + std::to_string(i) + ".";
    QfLayer& L = E->qf[i];
    memset(&L, 0, sizeof(L));
    L.cross = (i % c.qf_cross_freq) == 0;
    CGPT_TRY(mat(E, o + "qkv.w", 3 * Hq, Hq, &L.qkvw));
    CGPT_TRY(vec(E, o + "qkv.b", 3 * Hq, &L.qkvb));
    CGPT_TRY(mat(E, o + "ao.w", Hq, Hq, &L.aow));
    CGPT_TRY(vec(E, o + "ao.b", Hq, &L.aob));
    CGPT_TRY(vec(E, o + "aln.w", Hq, &L.alnw));
    CGPT_TRY(vec(E, o + "aln.b", Hq, &L.alnb));
    if (L.cross) {
      CGPT_TRY(mat(E, o + "cq.w", Hq, Hq, &L.cqw));
      CGPT_TRY(vec(E, o + "cq.b", Hq, &L.cqb));
      CGPT_TRY(mat(E, o + "co.w", Hq, Hq, &L.cow));
      CGPT_TRY(vec(E, o + "co.b", Hq, &L.cob));
      CGPT_TRY(vec(E, o + "cln.w", Hq, &L.clnw));
      CGPT_TRY(vec(E, o + "cln.b", Hq, &L.clnb));
    }
    CGPT_TRY(mat(E, o + "fi.w", c.qf_inter, Hq, &L.fiw));
    CGPT_TRY(vec(E, o + "fi.b", c.qf_inter, &L.fib));
    CGPT_TRY(mat(E, o + "fo.w", Hq, c.qf_inter, &L.fow));
    CGPT_TRY(vec(E, o + "fo.b", Hq, &L.fob));
    CGPT_TRY(vec(E, o + "fln.w", Hq, &L.flnw));
    CGPT_TRY(vec(E, o + "fln.b", Hq, &L.flnb));
  }
  CGPT_TRY(mat(E, "qf.ckv.w", static_cast<long long>(E->n_cross) * 2 * Hq, D, &E->ckv_w));
  CGPT_TRY(vec(E, "qf.ckv.b", static_cast<long long>(E->n_cross) * 2 * Hq, &E->ckv_b));
  CGPT_TRY(mat(E, "proj.w", Hl, Hq, &E->proj_w));
  CGPT_TRY(vec(E, "proj.b", Hl, &E->proj_b));
  CGPT_TRY(mat(E, "emb", c.llm_vocab, Hl, &E->emb));
  E->llm.resize(c.llm_layers);
  for (int i = 0; i < c.llm_layers; ++i) {
    const std::string o = "llm." + std::to_string(i) + ".";
    LlmLayer& L = E->llm[i];
    CGPT_TRY(mat(E, o + "qkv.w", 3 * Hl, Hl, &L.qkvw));
    CGPT_TRY(mat(E, o + "o.w", Hl, Hl, &L.ow));
    CGPT_TRY(mat(E, o + "gu.w", 2 * c.llm_inter, Hl, &L.guw));
    CGPT_TRY(mat(E, o + "down.w", Hl, c.llm_inter, &L.downw));
    CGPT_TRY(vec(E, o + "n1", Hl, &L.n1));
    CGPT_TRY(vec(E, o + "n2", Hl, &L.n2));
  }
  CGPT_TRY(vec(E, "llm.norm", Hl, &E->llm_norm));
  CGPT_TRY(mat(E, "llm.head", c.llm_vocab, Hl, &E->head_w));
  {
    auto it = E->w.find("rope.cos");
    CGPT_REQUIRE(it != E->w.end() && E->w.count("rope.sin"), "engine: rope.cos / rope.sin are not bound");
    CGPT_REQUIRE(it->second.dtype == CGPT_DT_F32 && it->second.cols == E->lhd / 2 &&
                     it->second.rows >= E->P + E->Tp_max + c.max_new_tokens,
                 "engine: rope tables must be f32 [>= %d, %d]", E->P + E->Tp_max + c.max_new_tokens, E->lhd / 2);
    E->rope_cos = static_cast<const float*>(it->second.p);
    E->rope_sin = static_cast<const float*>(E->w["rope.sin"].p);
  }
  E->resolved = true;
  return 0;
}

// ---------------------------------------------------------------- workspace
// One pass computes the size (base == nullptr) or assigns the pointers.
long long layout(Engine* E, int B, bool encoder_only, char* base, Buffers* out) {
  const cgpt_model_config& c = E->c;
  long long off = 0;
  auto take = [&](long long bytes) -> void* {
    void* p = base ? base + off : nullptr;
    off += align_up(bytes);
    return p;
  };
  const long long Mv = static_cast<long long>(B) * E->T, Mq = static_cast<long long>(B) * c.qf_queries,
                  Ml = static_cast<long long>(B) * E->Tp_max;
  const long long D = c.vit_dim, Hq = c.qf_hidden, Hl = c.llm_hidden;
  Buffers b;
  memset(&b, 0, sizeof(b));
  const long long KI = E->max_images;
  b.x_static = static_cast<float*>(take(KI * 3LL * c.img_size * c.img_size * 4));   // [max_images][3, S, S]
  b.dyn = take(KI * sizeof(NoiseDyn));
  b.counts = static_cast<long long*>(take(KI * 2LL * c.num_classes * 8));           // [max_images][2][num_classes]
  b.invalid = static_cast<int32_t*>(take(4));
  b.tail_label = static_cast<int32_t*>(take(KI * 3 * 4));
  b.tail_stats = static_cast<double*>(take(KI * 3 * 8));
  b.patches = take(static_cast<long long>(B) * E->Pn * 592 * 2);
  b.v_res = take(Mv * D * 4);
  b.v_xn = take(Mv * D * 2);
  b.v_qkv = take(Mv * 3 * D * 2);
  b.v_att = take(Mv * D * 2);
  b.v_h = take(Mv * c.vit_mlp * 2);
  b.v_out = take(Mv * D * 2);
  b.q_h = take(Mq * Hq * 2);
  b.q_tmp = take(Mq * Hq * 4);
  b.q_qkv = take(Mq * 3 * Hq * 2);
  b.q_ctx = take(Mq * Hq * 2);
  b.q_cq = take(Mq * Hq * 2);
  b.q_ckv = take(Mv * E->n_cross * 2 * Hq * 2);
  b.q_inter = take(Mq * c.qf_inter * 2);
  b.enc_out = take(Mq * Hl * 2);
  if (!encoder_only) {
    const long long rows = Ml > E->P ? Ml : E->P;   // the prefix pass (P rows) borrows these buffers
    b.l_res = take(rows * Hl * 4);
    b.l_xn = take(rows * Hl * 2);
    b.l_qkv = take(rows * 3 * Hl * 2);
    b.l_att = take(rows * Hl * 2);
    b.l_act = take(rows * c.llm_inter * 2);
    const long long cache = static_cast<long long>(c.llm_layers) * B * E->cache_rows * Hl * 2;
    b.kc = take(cache);
    b.vc = take(cache);
    const long long pre = static_cast<long long>(c.llm_layers) * (E->P > 0 ? E->P : 1) * Hl * 2;
    b.kp = take(pre);
    b.vp = take(pre);
    b.l_last = take(static_cast<long long>(B) * Hl * 2);
    b.l_res_last = take(static_cast<long long>(B) * Hl * 4);
    b.l_att_last = take(static_cast<long long>(B) * Hl * 2);
    b.l_logits = take(static_cast<long long>(B) * c.llm_vocab * 4);
    b.ids = static_cast<int32_t*>(take(static_cast<long long>(B) * c.max_new_tokens * 4));
    b.finished = static_cast<int32_t*>(take(B * 4LL));
    b.unfinished = static_cast<int32_t*>(take(4));
    b.next = static_cast<int32_t*>(take(B * 4LL));
    b.cur = static_cast<int32_t*>(take(B * 4LL));
    b.labels = static_cast<int32_t*>(take(B * 4LL));
    b.mcol = static_cast<float*>(take(B * 4LL));
    b.margin = static_cast<float*>(take(static_cast<long long>(B) * c.max_new_tokens * 4));
  }
  if (out) *out = b;
  return off;
}

void drop_graphs(Engine* E) {
  for (auto& g : E->graphs) {
    if (g.g0) cudaGraphExecDestroy(g.g0);
    for (auto s : g.steps)
      if (s) cudaGraphExecDestroy(s);
  }
  E->graphs.clear();
}

// ---------------------------------------------------------------- towers
// A6-A8: patches -> ln_vision(ViT features) (eva_vit.py:204-210,332-349; base_model.py:281-287)
int vit_forward(Engine* E, const void* patches, int B, void* out, cudaStream_t s) {
  const cgpt_model_config& c = E->c;
  const Buffers& b = E->b;
  const int T = E->T, Pn = E->Pn, D = c.vit_dim;
  const int M = B * T;
  float* res = static_cast<float*>(b.v_res);
  // conv14 as GEMM; the epilogue adds bias + pos_embed[1+p] and scatters row b*Pn+p -> b*T+1+p
  CGPT_TRY(gemm(patches, 592, E->patch_w, B * Pn, D, 592,
                Epi(res, D, CGPT_DT_F32).bias(E->patch_b).row_add(E->pos, D, 1).remap(Pn, T, 1), s));
  CGPT_TRY(set_rows_f32(res, static_cast<long long>(T) * D, E->cls_pos, D, B, s));
  const float scale = 1.0f / sqrtf(static_cast<float>(E->vhd));
  // head-major q / k / v ([3][B][H][T][hd], written by the QKV GEMM's scatter epilogue) + the pipelined tcgen05
  // attention kernels whenever the shape allows (224 px: T = 257, one key tile + cls; 448 px: T = 1025, multi-tile);
  // row-major + the generic dispatcher otherwise (head dims <= 64: the small test models)
  cgpt_attn_args hm;
  memset(&hm, 0, sizeof(hm));
  const long long MD = static_cast<long long>(M) * D;
  hm.q = b.v_qkv; hm.k = bf(b.v_qkv) + MD; hm.v = bf(b.v_qkv) + 2 * MD;
  hm.q_rows_per_batch = T; hm.kv_rows_per_batch = T; hm.o = b.v_att; hm.ldo = D;
  hm.B = B; hm.H = c.vit_heads; hm.Tq = T; hm.Tk = T; hm.head_dim = E->vhd; hm.scale = scale; hm.head_major = 1;
  static const bool no_hm = getenv("CGPT_VIT_ROW_MAJOR") != nullptr;   // A/B switch: the round-1 layout and kernel
  const bool use_hm = !no_hm && (attn_vit_supported(&hm) || attn_long_supported(&hm));
  for (int i = 0; i < c.vit_depth; ++i) {
    const VitLayer& L = E->vit[i];
    CGPT_TRY(norm_rows(res, D, CGPT_DT_F32, L.ln1w, L.ln1b, c.vit_eps, M, D, b.v_xn, D, CGPT_DT_BF16, 0, 0, 0, 0, s));
    if (use_hm) {
      CGPT_TRY(gemm(b.v_xn, D, L.qkvw, M, 3 * D, D,
                    Epi(b.v_qkv, 3 * D, CGPT_DT_BF16).bias(L.qkvb).headmajor(T, c.vit_heads, E->vhd), s));
      CGPT_TRY(attention(&hm, s));
    } else {
      CGPT_TRY(gemm(b.v_xn, D, L.qkvw, M, 3 * D, D, Epi(b.v_qkv, 3 * D, CGPT_DT_BF16).bias(L.qkvb), s));
      CGPT_TRY(attn(b.v_qkv, 3 * D, T, bf(b.v_qkv) + D, bf(b.v_qkv) + 2 * D, 3 * D, T, b.v_att, D, B, c.vit_heads, T, T,
                    E->vhd, scale, 0, 0, s));
    }
    CGPT_TRY(gemm(b.v_att, D, L.projw, M, D, D,
                  Epi(res, D, CGPT_DT_F32).bias(L.projb).resid(res, D, CGPT_DT_F32), s));
    CGPT_TRY(norm_rows(res, D, CGPT_DT_F32, L.ln2w, L.ln2b, c.vit_eps, M, D, b.v_xn, D, CGPT_DT_BF16, 0, 0, 0, 0, s));
    CGPT_TRY(gemm(b.v_xn, D, L.fc1w, M, c.vit_mlp, D,
                  Epi(b.v_h, c.vit_mlp, CGPT_DT_BF16).bias(L.fc1b).act(CGPT_ACT_GELU), s));
    CGPT_TRY(gemm(b.v_h, c.vit_mlp, L.fc2w, M, D, c.vit_mlp,
                  Epi(res, D, CGPT_DT_F32).bias(L.fc2b).resid(res, D, CGPT_DT_F32), s));
  }
  CGPT_TRY(norm_rows(res, D, CGPT_DT_F32, E->lnv_w, E->lnv_b, c.ln_vision_eps, M, D, out, D, CGPT_DT_BF16, 0, 0, 0, 0, s));
  return 0;
}

// A9: Q-Former on 32 queries, cross K/V of all cross layers from ONE GEMM (Qformer.py:402-484)
int qformer_forward(Engine* E, const void* image_embeds, int B, void* out_h, cudaStream_t s) {
  const cgpt_model_config& c = E->c;
  const Buffers& b = E->b;
  const int T = E->T, nq = c.qf_queries, Hd = c.qf_hidden, D = c.vit_dim;
  const int M = B * nq;
  const long long ldckv = static_cast<long long>(E->n_cross) * 2 * Hd;
  void* h = out_h;
  CGPT_TRY(gather_rows(E->qf_q0, Hd, nullptr, nq, M, Hd, h, Hd, CGPT_DT_BF16, 0, 0, 0, nq, s));
  CGPT_TRY(gemm(image_embeds, D, E->ckv_w, B * T, static_cast<int>(ldckv), D,
                Epi(b.q_ckv, ldckv, CGPT_DT_BF16).bias(E->ckv_b), s));
  const float scale = 1.0f / sqrtf(static_cast<float>(E->qhd));
  int ci = 0;
  for (int i = 0; i < c.qf_layers; ++i) {
    const QfLayer& L = E->qf[i];
    CGPT_TRY(gemm(h, Hd, L.qkvw, M, 3 * Hd, Hd, Epi(b.q_qkv, 3 * Hd, CGPT_DT_BF16).bias(L.qkvb), s));
    CGPT_TRY(attn(b.q_qkv, 3 * Hd, nq, bf(b.q_qkv) + Hd, bf(b.q_qkv) + 2 * Hd, 3 * Hd, nq, b.q_ctx, Hd, B, c.qf_heads,
                  nq, nq, E->qhd, scale, 0, 0, s));
    CGPT_TRY(gemm(b.q_ctx, Hd, L.aow, M, Hd, Hd,
                  Epi(b.q_tmp, Hd, CGPT_DT_F32).bias(L.aob).resid(h, Hd, CGPT_DT_BF16), s));
    CGPT_TRY(norm_rows(b.q_tmp, Hd, CGPT_DT_F32, L.alnw, L.alnb, c.qf_eps, M, Hd, h, Hd, CGPT_DT_BF16, 0, 0, 0, 0, s));
    if (L.cross) {
      CGPT_TRY(gemm(h, Hd, L.cqw, M, Hd, Hd, Epi(b.q_cq, Hd, CGPT_DT_BF16).bias(L.cqb), s));
      const long long kcol = static_cast<long long>(ci) * 2 * Hd;
      CGPT_TRY(attn(b.q_cq, Hd, nq, bf(b.q_ckv) + kcol, bf(b.q_ckv) + kcol + Hd, ldckv, T, b.q_ctx, Hd, B, c.qf_heads,
                    nq, T, E->qhd, scale, 0, 0, s));
      CGPT_TRY(gemm(b.q_ctx, Hd, L.cow, M, Hd, Hd,
                    Epi(b.q_tmp, Hd, CGPT_DT_F32).bias(L.cob).resid(h, Hd, CGPT_DT_BF16), s));
      CGPT_TRY(norm_rows(b.q_tmp, Hd, CGPT_DT_F32, L.clnw, L.clnb, c.qf_eps, M, Hd, h, Hd, CGPT_DT_BF16, 0, 0, 0, 0, s));
      ++ci;
    }
    CGPT_TRY(gemm(h, Hd, L.fiw, M, c.qf_inter, Hd,
                  Epi(b.q_inter, c.qf_inter, CGPT_DT_BF16).bias(L.fib).act(CGPT_ACT_GELU), s));
    CGPT_TRY(gemm(b.q_inter, c.qf_inter, L.fow, M, Hd, c.qf_inter,
                  Epi(b.q_tmp, Hd, CGPT_DT_F32).bias(L.fob).resid(h, Hd, CGPT_DT_BF16), s));
    CGPT_TRY(norm_rows(b.q_tmp, Hd, CGPT_DT_F32, L.flnw, L.flnb, c.qf_eps, M, Hd, h, Hd, CGPT_DT_BF16, 0, 0, 0, 0, s));
  }
  return 0;
}

// ---------------------------------------------------------------- Llama
// HF LlamaDecoderLayer x L on `rows` = B*T rows; K/V appended to the cache at row cache_row0 of each sample
// last_res (nullable): the caller only needs the LAST position of every sample after the stack (prefill: the
// position that predicts the first new token).  The last layer then still projects K/V for all rows, but its
// o_proj, MLP and residual run on B rows instead of B*T (same per-row arithmetic);
// the compact residual [B, hidden] fp32 lands in last_res.
int llm_layers(Engine* E, int rows, int T, int B, void* res, void* xn, void* qkv, void* att, void* act, void* kc,
               void* vc, long long layer_stride, int pos0, int cache_row0, int cache_rows, int decode,
               cudaStream_t s, void* last_res = nullptr, void* last_att = nullptr) {
  const cgpt_model_config& c = E->c;
  const int Hd = c.llm_hidden;
  const float scale = 1.0f / sqrtf(static_cast<float>(E->lhd));
  static const bool no_fused_rope = getenv("CGPT_NO_FUSED_ROPE") != nullptr;   // A/B switch
  const bool fused_rope = E->lhd == 128 && !no_fused_rope;
  for (int i = 0; i < c.llm_layers; ++i) {
    const LlmLayer& L = E->llm[i];
    void* kci = bf(kc) + i * layer_stride;
    void* vci = bf(vc) + i * layer_stride;
    CGPT_TRY(norm_rows(res, Hd, CGPT_DT_F32, L.n1, nullptr, c.llm_rms_eps, rows, Hd, xn, Hd, CGPT_DT_BF16, 1, 0, 0, 0, s));
    if (fused_rope) {
      // rotary embedding + KV-cache append inside the QKV GEMM's epilogue (128-wide heads)
      cgpt_gemm_rope r;
      r.T = T; r.heads = c.llm_heads; r.pos0 = pos0;
      r.cos_table = E->rope_cos; r.sin_table = E->rope_sin;
      r.kcache = kci; r.vcache = vci; r.ld_cache = Hd;
      r.cache_rows_per_batch = cache_rows; r.cache_row0 = cache_row0;
      CGPT_TRY(gemm(xn, Hd, L.qkvw, rows, 3 * Hd, Hd, Epi(qkv, 3 * Hd, CGPT_DT_BF16).rope(&r), s));
    } else {
      CGPT_TRY(gemm(xn, Hd, L.qkvw, rows, 3 * Hd, Hd, Epi(qkv, 3 * Hd, CGPT_DT_BF16), s));
      CGPT_TRY(rope_split(qkv, 3 * Hd, rows, T, c.llm_heads, E->lhd, pos0, E->rope_cos, E->rope_sin, kci, vci, Hd,
                          cache_rows, cache_row0, s));
    }
    CGPT_TRY(attn(qkv, 3 * Hd, T, kci, vci, Hd, cache_rows, att, Hd, B, c.llm_heads, T, cache_row0 + T, E->lhd, scale,
                  1, decode, s));
    if (last_res != nullptr && i == c.llm_layers - 1 && T > 1) {
      CGPT_TRY(gather_last_rows(last_res, res, T, B, static_cast<long long>(Hd) * 4, s));
      CGPT_TRY(gather_last_rows(last_att, att, T, B, static_cast<long long>(Hd) * 2, s));
      CGPT_TRY(gemm(last_att, Hd, L.ow, B, Hd, Hd, Epi(last_res, Hd, CGPT_DT_F32).resid(last_res, Hd, CGPT_DT_F32), s));
      CGPT_TRY(norm_rows(last_res, Hd, CGPT_DT_F32, L.n2, nullptr, c.llm_rms_eps, B, Hd, xn, Hd, CGPT_DT_BF16, 1, 0, 0, 0, s));
      CGPT_TRY(gemm(xn, Hd, L.guw, B, 2 * c.llm_inter, Hd, Epi(act, c.llm_inter, CGPT_DT_BF16).act(CGPT_ACT_SWIGLU), s));
      CGPT_TRY(gemm(act, c.llm_inter, L.downw, B, Hd, c.llm_inter,
                    Epi(last_res, Hd, CGPT_DT_F32).resid(last_res, Hd, CGPT_DT_F32), s));
      return 0;
    }
    CGPT_TRY(gemm(att, Hd, L.ow, rows, Hd, Hd, Epi(res, Hd, CGPT_DT_F32).resid(res, Hd, CGPT_DT_F32), s));
    CGPT_TRY(norm_rows(res, Hd, CGPT_DT_F32, L.n2, nullptr, c.llm_rms_eps, rows, Hd, xn, Hd, CGPT_DT_BF16, 1, 0, 0, 0, s));
    CGPT_TRY(gemm(xn, Hd, L.guw, rows, 2 * c.llm_inter, Hd, Epi(act, c.llm_inter, CGPT_DT_BF16).act(CGPT_ACT_SWIGLU), s));
    CGPT_TRY(gemm(act, c.llm_inter, L.downw, rows, Hd, c.llm_inter,
                  Epi(res, Hd, CGPT_DT_F32).resid(res, Hd, CGPT_DT_F32), s));
  }
  return 0;
}

// K/V of the batch-invariant prompt prefix ("<s>[INST] <Img>", the tokens BEFORE the image), computed once:
// exact under causal attention; then replicated into rows [0, P) of every sample's cache
int build_prefix(Engine* E, cudaStream_t s) {
  const cgpt_model_config& c = E->c;
  const Buffers& b = E->b;
  const int P = E->P, Hd = c.llm_hidden;
  const long long cache_bytes = static_cast<long long>(c.llm_layers) * E->ws_B * E->cache_rows * Hd * 2;
  CGPT_CHECK_CUDA(cudaMemsetAsync(b.kc, 0, cache_bytes, s));
  CGPT_CHECK_CUDA(cudaMemsetAsync(b.vc, 0, cache_bytes, s));
  if (P == 0) return 0;
  CGPT_TRY(gather_rows(E->emb, Hd, E->prefix_ids, P, P, Hd, b.l_res, Hd, CGPT_DT_F32, 0, 0, 0, c.llm_vocab, s));
  CGPT_TRY(llm_layers(E, P, P, 1, b.l_res, b.l_xn, b.l_qkv, b.l_att, b.l_act, b.kp, b.vp,
                      static_cast<long long>(P) * Hd, 0, 0, P, 0, s));
  const long long layer_stride = static_cast<long long>(E->ws_B) * E->cache_rows * Hd;
  for (int i = 0; i < c.llm_layers; ++i) {
    CGPT_TRY(gather_rows(bf(b.kp) + static_cast<long long>(i) * P * Hd, Hd, nullptr, P, E->ws_B * P, Hd,
                         bf(b.kc) + i * layer_stride, Hd, CGPT_DT_BF16, P, E->cache_rows, 0, P, s));
    CGPT_TRY(gather_rows(bf(b.vp) + static_cast<long long>(i) * P * Hd, Hd, nullptr, P, E->ws_B * P, Hd,
                         bf(b.vc) + i * layer_stride, Hd, CGPT_DT_BF16, P, E->cache_rows, 0, P, s));
  }
  return 0;
}

// lm_head on the last hidden state, greedy pick with HF min_length / EOS / pad bookkeeping
int head_and_pick(Engine* E, int B, int t, cudaStream_t s) {
  const cgpt_model_config& c = E->c;
  const Buffers& b = E->b;
  const int Hd = c.llm_hidden, mn = c.max_new_tokens;
  CGPT_TRY(gemm(b.l_last, Hd, E->head_w, B, c.llm_vocab, Hd, Epi(b.l_logits, c.llm_vocab, CGPT_DT_F32), s));
  CGPT_TRY(argmax_rows(static_cast<const float*>(b.l_logits), B, c.llm_vocab, c.llm_vocab,
                       t < c.min_length ? c.eos_id : -1, b.next, b.mcol, s));
  CGPT_TRY(copy_col_u32(b.margin + t, mn, b.mcol, 1, B, s));
  CGPT_CHECK_CUDA(cudaMemsetAsync(b.unfinished, 0, 4, s));
  CGPT_TRY(greedy_step(b.next, B, b.finished, b.ids, mn, t, c.eos_id, c.pad_id, b.unfinished, s));
  return 0;
}

// A10-A12: llama_proj + prompt assembly + prefill + first greedy token (minigpt4.py:141; minigpt_base.py:75-89,414-427)
int llm_prefill_first(Engine* E, const void* qf_out, int B, cudaStream_t s) {
  const cgpt_model_config& c = E->c;
  const Buffers& b = E->b;
  const int nq = c.qf_queries, Tp = E->Tp, P = E->P, Hd = c.llm_hidden, ns = E->ns, mn = c.max_new_tokens;
  const int M = B * Tp;
  const long long layer_stride = static_cast<long long>(E->ws_B) * E->cache_rows * Hd;
  // llama_proj output lands in rows [b*Tp, b*Tp+nq); the question embeddings are broadcast behind it
  CGPT_TRY(gemm(qf_out, c.qf_hidden, E->proj_w, B * nq, Hd, c.qf_hidden,
                Epi(b.l_res, Hd, CGPT_DT_F32).bias(E->proj_b).remap(nq, Tp, 0), s));
  if (ns > 0)
    CGPT_TRY(gather_rows(E->emb, Hd, E->suffix_ids, ns, B * ns, Hd, b.l_res, Hd, CGPT_DT_F32, ns, Tp, nq, c.llm_vocab, s));
  // the last layer only needs the last prompt position of every sample (CGPT_NO_LAST_PRUNE=1: A/B switch)
  static const bool no_prune = getenv("CGPT_NO_LAST_PRUNE") != nullptr;
  const bool prune = Tp > 1 && !no_prune;
  CGPT_TRY(llm_layers(E, M, Tp, B, b.l_res, b.l_xn, b.l_qkv, b.l_att, b.l_act, b.kc, b.vc, layer_stride, P, P,
                      E->cache_rows, 0, s, prune ? b.l_res_last : nullptr, b.l_att_last));
  CGPT_TRY(fill_u32(b.ids, static_cast<long long>(B) * mn, static_cast<uint32_t>(c.pad_id), s));
  CGPT_CHECK_CUDA(cudaMemsetAsync(b.finished, 0, B * 4LL, s));
  CGPT_TRY(fill_u32(b.margin, static_cast<long long>(B) * mn, 0x7f800000u /* +inf */, s));
  if (prune)
    CGPT_TRY(norm_rows(b.l_res_last, Hd, CGPT_DT_F32, E->llm_norm, nullptr, c.llm_rms_eps, B, Hd, b.l_last, Hd,
                       CGPT_DT_BF16, 1, 0, 0, 0, s));
  else
    CGPT_TRY(norm_rows(b.l_res, Hd, CGPT_DT_F32, E->llm_norm, nullptr, c.llm_rms_eps, B, Hd, b.l_last, Hd, CGPT_DT_BF16, 1,
                       1, Tp, Tp - 1, s));
  return head_and_pick(E, B, 0, s);
}

// decode step t >= 1: embed token t-1, one row per sample against the KV cache, pick token t.
// The single-row buffers alias the first B rows of the prefill buffers (prefill is complete by then).
int llm_decode_step(Engine* E, int B, int t, cudaStream_t s) {
  const cgpt_model_config& c = E->c;
  const Buffers& b = E->b;
  const int Hd = c.llm_hidden, mn = c.max_new_tokens;
  const long long layer_stride = static_cast<long long>(E->ws_B) * E->cache_rows * Hd;
  const int row = E->P + E->Tp + t - 1;
  CGPT_TRY(copy_col_u32(b.cur, 1, b.ids + (t - 1), mn, B, s));
  CGPT_TRY(gather_rows(E->emb, Hd, b.cur, B, B, Hd, b.l_res, Hd, CGPT_DT_F32, 0, 0, 0, c.llm_vocab, s));
  const int decode = ((E->lhd == 32 || E->lhd == 64 || E->lhd == 128) && c.llm_heads % 4 == 0) ? 1 : 0;
  CGPT_TRY(llm_layers(E, B, 1, B, b.l_res, b.l_xn, b.l_qkv, b.l_att, b.l_act, b.kc, b.vc, layer_stride, row, row,
                      E->cache_rows, decode, s));
  CGPT_TRY(norm_rows(b.l_res, Hd, CGPT_DT_F32, E->llm_norm, nullptr, c.llm_rms_eps, B, Hd, b.l_last, Hd, CGPT_DT_BF16, 1,
                     0, 0, 0, s));
  return head_and_pick(E, B, t, s);
}

int all_finished(Engine* E, cudaStream_t s, bool* done) {
  *done = false;
  if (!E->c.early_exit) return 0;
  CGPT_CHECK_CUDA(cudaMemcpyAsync(E->h_i32, E->b.unfinished, 4, cudaMemcpyDeviceToHost, s));
  CGPT_CHECK_CUDA(cudaStreamSynchronize(s));
  *done = E->h_i32[0] == 0;
  return 0;
}

int llm_generate(Engine* E, const void* qf_out, int B, cudaStream_t s) {
  CGPT_TRY(llm_prefill_first(E, qf_out, B, s));
  int steps = 1;
  for (int t = 1; t < E->c.max_new_tokens; ++t) {
    bool done;
    CGPT_TRY(all_finished(E, s, &done));
    if (done) break;
    CGPT_TRY(llm_decode_step(E, B, t, s));
    steps = t + 1;
  }
  E->last_steps = steps;
  return 0;
}

int labels_of_ids(Engine* E, int B, int32_t* labels, cudaStream_t s) {
  CGPT_REQUIRE(E->table_keys != nullptr, "engine: answer table is not set (cgpt_set_answer_table)");
  return answer_labels(E->b.ids, B, E->c.max_new_tokens, E->c.max_new_tokens, E->c.eos_id, E->table_keys,
                       E->table_vals, E->table_cap, E->c.num_classes - 1, labels, s);
}

int check_ready(Engine* E, int B, bool need_llm) {
  CGPT_REQUIRE(E != nullptr, "engine: null handle");
  CGPT_REQUIRE(E->ws != nullptr, "engine: no workspace bound (cgpt_bind_workspace)");
  CGPT_REQUIRE(B > 0 && B <= E->ws_B, "engine: batch %d exceeds the workspace batch %d", B, E->ws_B);
  CGPT_REQUIRE(!need_llm || !E->ws_encoder_only, "engine: the workspace was bound encoder-only");
  return 0;
}

// ---------------------------------------------------------------- one batch: eager and graph replay
int stage0_eager(Engine* E, const float* x, const float* eps, const cgpt_noise_spec* n, uint64_t first, int B,
                 cudaStream_t s) {
  const Buffers& b = E->b;
  CGPT_TRY(noise_patchify(x, eps, n->seed, n->stream_id, first, B, n->sigma, n->mean, n->std, n->noise_space,
                          n->noise_kind, E->c.img_size, b.patches, 592, nullptr, s));
  CGPT_TRY(vit_forward(E, b.patches, B, b.v_out, s));
  CGPT_TRY(qformer_forward(E, b.v_out, B, b.q_h, s));
  return 0;
}

int capture_begin(Engine* E) {
  CGPT_CHECK_CUDA(cudaStreamBeginCapture(E->cap_stream, cudaStreamCaptureModeRelaxed));
  return 0;
}
int capture_end(Engine* E, int body_rc, cudaGraphExec_t* exec) {
  cudaGraph_t g = nullptr;
  cudaError_t e = cudaStreamEndCapture(E->cap_stream, &g);
  if (body_rc != 0) {
    if (g) cudaGraphDestroy(g);
    return body_rc;
  }
  CGPT_CHECK_CUDA(e);
  cudaError_t ei = cudaGraphInstantiate(exec, g, 0);
  cudaGraphDestroy(g);
  CGPT_CHECK_CUDA(ei);
  return 0;
}

// K1 for a pass over K images: segment k = B / K draws of the image at x_static[k], noise parameters from dyn[k]
int noise_segments(Engine* E, const cgpt_noise_spec* n, int B, int K, cudaStream_t s) {
  const Buffers& b = E->b;
  const int per = B / K;
  const long long img_elems = 3LL * E->c.img_size * E->c.img_size;
  for (int k = 0; k < K; ++k)
    CGPT_TRY(noise_patchify(b.x_static + k * img_elems, nullptr, 0, 0, 0, per, 0.f, n->mean, n->std, n->noise_space,
                            n->noise_kind, E->c.img_size, bf(b.patches) + static_cast<long long>(k) * per * E->Pn * 592,
                            592, static_cast<const NoiseDyn*>(b.dyn) + k, s));
  return 0;
}

int get_graphs(Engine* E, const cgpt_noise_spec* n, int B, int K, cudaStream_t s, GraphSet** out) {
  for (auto& g : E->graphs)
    if (g.B == B && g.K == K && g.ns == E->ns && g.suffix == E->suffix_ids && g.space == n->noise_space && g.kind == n->noise_kind && !memcmp(g.mean, n->mean, 12) &&
        !memcmp(g.std, n->std, 12)) {
      *out = &g;
      return 0;
    }
  const Buffers& b = E->b;
  // warm-up outside capture (first-use cudaFuncSetAttribute calls), on the caller's stream
  CGPT_TRY(noise_segments(E, n, B, K, s));
  CGPT_TRY(vit_forward(E, b.patches, B, b.v_out, s));
  CGPT_TRY(qformer_forward(E, b.v_out, B, b.q_h, s));
  CGPT_TRY(llm_prefill_first(E, b.q_h, B, s));
  for (int t = 1; t < E->c.max_new_tokens; ++t) CGPT_TRY(llm_decode_step(E, B, t, s));
  CGPT_CHECK_CUDA(cudaStreamSynchronize(s));

  GraphSet gs;
  gs.B = B; gs.K = K; gs.space = n->noise_space; gs.kind = n->noise_kind;
  gs.ns = E->ns; gs.suffix = E->suffix_ids;
  memcpy(gs.mean, n->mean, 12);
  memcpy(gs.std, n->std, 12);
  cudaStream_t cs = E->cap_stream;
  const long long c_start = total_launch_count();
  long long c0 = c_start;
  CGPT_TRY(capture_begin(E));
  int rc = noise_segments(E, n, B, K, cs);
  if (!rc) rc = vit_forward(E, b.patches, B, b.v_out, cs);
  if (!rc) rc = qformer_forward(E, b.v_out, B, b.q_h, cs);
  if (!rc) rc = llm_prefill_first(E, b.q_h, B, cs);
  CGPT_TRY(capture_end(E, rc, &gs.g0));
  gs.nodes.push_back(static_cast<int>(total_launch_count() - c0));
  for (int t = 1; t < E->c.max_new_tokens; ++t) {
    c0 = total_launch_count();
    CGPT_TRY(capture_begin(E));
    rc = llm_decode_step(E, B, t, cs);
    cudaGraphExec_t ge = nullptr;
    CGPT_TRY(capture_end(E, rc, &ge));
    gs.steps.push_back(ge);
    gs.nodes.push_back(static_cast<int>(total_launch_count() - c0));
  }
  // captured launches did not run: take them out of the launch counter (replays add them back)
  count_launch(-static_cast<int>(total_launch_count() - c_start));
  E->graphs.push_back(gs);
  *out = &E->graphs.back();
  return 0;
}

int noisy_labels(Engine* E, const float* x_dev, const cgpt_noise_spec* n, uint64_t first, int B, int32_t* labels,
                 cudaStream_t s) {
  CGPT_TRY(check_ready(E, B, true));
  const Buffers& b = E->b;
  const long long img_elems = 3LL * E->c.img_size * E->c.img_size;
  if (E->c.use_graphs && n->eps == nullptr) {
    if (x_dev != b.x_static)
      CGPT_CHECK_CUDA(cudaMemcpyAsync(b.x_static, x_dev, img_elems * 4, cudaMemcpyDefault, s));
    NoiseDyn d;
    d.seed = n->seed; d.first_sample = first; d.stream_id = n->stream_id; d.sigma = n->sigma;
    // pageable source: staged before the call returns, so the stack bytes can die at once
    CGPT_CHECK_CUDA(cudaMemcpyAsync(b.dyn, &d, sizeof(d), cudaMemcpyHostToDevice, s));
    GraphSet* g = nullptr;
    CGPT_TRY(get_graphs(E, n, B, 1, s, &g));
    CGPT_CHECK_CUDA(cudaGraphLaunch(g->g0, s));
    count_launch(g->nodes[0]);
    int steps = 1;
    for (int t = 1; t < E->c.max_new_tokens; ++t) {
      bool done;
      CGPT_TRY(all_finished(E, s, &done));
      if (done) break;
      CGPT_CHECK_CUDA(cudaGraphLaunch(g->steps[t - 1], s));
      count_launch(g->nodes[t]);
      steps = t + 1;
    }
    E->last_steps = steps;
  } else {
    const float* eps = n->eps ? n->eps + static_cast<long long>(first) * img_elems : nullptr;
    CGPT_TRY(stage0_eager(E, x_dev, eps, n, first, B, s));
    CGPT_TRY(llm_generate(E, b.q_h, B, s));
  }
  return labels_of_ids(E, B, labels, s);
}

// One pass over K images (already staged in x_static[0..K)): `per` draws of each, global sample indices
// [first, first + per) of every image, image k drawn from Philox stream stream_id + k.  labels: [K * per], image-major.
int noisy_labels_images(Engine* E, const cgpt_noise_spec* n, uint64_t first, int K, int per, int32_t* labels,
                        cudaStream_t s) {
  const int B = K * per;
  CGPT_TRY(check_ready(E, B, true));
  CGPT_REQUIRE(K >= 1 && K <= E->max_images, "engine: %d images per pass, the workspace serves %d (option max_images)", K,
               E->max_images);
  CGPT_REQUIRE(n->eps == nullptr, "engine: injected noise is a single-image (parity) feature");
  const Buffers& b = E->b;
  std::vector<NoiseDyn> d(K);
  for (int k = 0; k < K; ++k) {
    d[k].seed = n->seed; d[k].first_sample = first; d[k].stream_id = n->stream_id + static_cast<uint32_t>(k);
    d[k].sigma = n->sigma;
  }
  // pageable source: staged before the call returns, so the host vector can die at once
  CGPT_CHECK_CUDA(cudaMemcpyAsync(b.dyn, d.data(), K * sizeof(NoiseDyn), cudaMemcpyHostToDevice, s));
  if (E->c.use_graphs) {
    GraphSet* g = nullptr;
    CGPT_TRY(get_graphs(E, n, B, K, s, &g));
    CGPT_CHECK_CUDA(cudaGraphLaunch(g->g0, s));
    count_launch(g->nodes[0]);
    int steps = 1;
    for (int t = 1; t < E->c.max_new_tokens; ++t) {
      bool done;
      CGPT_TRY(all_finished(E, s, &done));
      if (done) break;
      CGPT_CHECK_CUDA(cudaGraphLaunch(g->steps[t - 1], s));
      count_launch(g->nodes[t]);
      steps = t + 1;
    }
    E->last_steps = steps;
  } else {
    CGPT_TRY(noise_segments(E, n, B, K, s));
    CGPT_TRY(vit_forward(E, b.patches, B, b.v_out, s));
    CGPT_TRY(qformer_forward(E, b.v_out, B, b.q_h, s));
    CGPT_TRY(llm_generate(E, b.q_h, B, s));
  }
  return labels_of_ids(E, B, labels, s);
}

// Smooth._sample_noise (smoothing.py:81-99) on this rank's slice of [base, base + num)
int sample_noise(Engine* E, const float* x, const cgpt_noise_spec* n, long long base, long long num, int batch_size,
                 long long split, int rank, int world, void* comm, long long* counts, int32_t* invalid,
                 cudaStream_t s) {
  CGPT_REQUIRE(E != nullptr && x != nullptr && n != nullptr && counts != nullptr, "sample_noise: null argument");
  CGPT_REQUIRE(num >= 0 && batch_size > 0 && world > 0 && rank >= 0 && rank < world,
               "sample_noise: bad sizes num=%lld batch_size=%d rank=%d world=%d", num, batch_size, rank, world);
  CGPT_REQUIRE(split < 0 || split <= num, "sample_noise: split %lld outside [0, %lld]", split, num);
  CGPT_TRY(check_ready(E, batch_size < E->ws_B ? batch_size : E->ws_B, true));
  if (batch_size > E->ws_B) batch_size = E->ws_B;
  const Buffers& b = E->b;
  const int nc = E->c.num_classes;
  const int nvec = split >= 0 ? 2 : 1;
  CGPT_CHECK_CUDA(cudaMemsetAsync(counts, 0, static_cast<size_t>(nvec) * nc * 8, s));
  int32_t* inv = invalid ? invalid : b.invalid;
  CGPT_CHECK_CUDA(cudaMemsetAsync(inv, 0, 4, s));
  // x -> static device buffer (host or device source)
  const long long img_bytes = 3LL * E->c.img_size * E->c.img_size * 4;
  if (x != b.x_static) CGPT_CHECK_CUDA(cudaMemcpyAsync(b.x_static, x, img_bytes, cudaMemcpyDefault, s));
  const long long lo = base + (num * rank) / world, hi = base + (num * (rank + 1)) / world;
  const long long boundary = split >= 0 ? base + split : -1;
  const long long n_chunks = hi > lo ? (hi - lo + batch_size - 1) / batch_size : 0;
  long long first = lo;
  for (long long ci = 0; ci < n_chunks; ++ci) {
    // balanced chunks (each <= batch_size): no tiny tail batch
    const int bs = static_cast<int>((hi - first + (n_chunks - ci) - 1) / (n_chunks - ci));
    CGPT_TRY(noisy_labels(E, b.x_static, n, static_cast<uint64_t>(first), bs, b.labels, s));
    if (boundary < 0 || first + bs <= boundary) {
      CGPT_TRY(label_hist(b.labels, bs, nc, counts, inv, s));
    } else if (first >= boundary) {
      CGPT_TRY(label_hist(b.labels, bs, nc, counts + nc, inv, s));
    } else {
      const int k = static_cast<int>(boundary - first);
      CGPT_TRY(label_hist(b.labels, k, nc, counts, inv, s));
      CGPT_TRY(label_hist(b.labels + k, bs - k, nc, counts + nc, inv, s));
    }
    first += bs;
  }
  if (world > 1 && comm != nullptr)
    CGPT_TRY(cgpt_allreduce_counts(reinterpret_cast<int64_t*>(counts), nvec * nc, comm, s));
  return 0;
}

// ---------------------------------------------------------------- NCCL (resolved at run time)
struct Id128 { char internal[128]; };   // ncclUniqueId (passed by value)
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, Id128, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;

int load_nccl() {
  if (g_nccl.lib) return 0;
  void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);   // the copy torch already loaded, if any
  if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW);
  if (!lib) lib = dlopen("libnccl.so", RTLD_NOW);
  CGPT_REQUIRE(lib != nullptr, "NCCL: cannot dlopen libnccl.so.2 (%s)", dlerror());
  g_nccl.GetUniqueId = reinterpret_cast<decltype(g_nccl.GetUniqueId)>(dlsym(lib, "ncclGetUniqueId"));
  g_nccl.CommInitRank = reinterpret_cast<decltype(g_nccl.CommInitRank)>(dlsym(lib, "ncclCommInitRank"));
  g_nccl.AllReduce = reinterpret_cast<decltype(g_nccl.AllReduce)>(dlsym(lib, "ncclAllReduce"));
  g_nccl.CommDestroy = reinterpret_cast<decltype(g_nccl.CommDestroy)>(dlsym(lib, "ncclCommDestroy"));
  g_nccl.GetErrorString = reinterpret_cast<decltype(g_nccl.GetErrorString)>(dlsym(lib, "ncclGetErrorString"));
  CGPT_REQUIRE(g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.AllReduce && g_nccl.CommDestroy,
               "NCCL: missing symbols in libnccl");
  g_nccl.lib = lib;
  return 0;
}
#define CGPT_CHECK_NCCL(expr)                                                                          \
  do {                                                                                                 \
    int _r = (expr);                                                                                   \
    if (_r != 0) {                                                                                     \
      set_last_error("NCCL error %d (%s) in %s", _r,                                                   \
                     g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : "?", #expr);                  \
      return -3;                                                                                       \
    }                                                                                                  \
  } while (0)

}  // namespace
}  // namespace cgpt

using namespace cgpt;

extern "C" {

int cgpt_create(const cgpt_model_config* cfg, cgpt_handle* out) {
  CGPT_REQUIRE(cfg != nullptr && out != nullptr, "cgpt_create: null argument");
  const cgpt_model_config& c = *cfg;
  CGPT_REQUIRE(c.img_size > 0 && c.img_size % 14 == 0, "cgpt_create: img_size %d must be a multiple of 14", c.img_size);
  CGPT_REQUIRE(c.vit_dim > 0 && c.vit_heads > 0 && c.vit_dim % c.vit_heads == 0 && c.vit_depth > 0 && c.vit_mlp > 0,
               "cgpt_create: bad ViT shape");
  CGPT_REQUIRE(c.qf_hidden > 0 && c.qf_heads > 0 && c.qf_hidden % c.qf_heads == 0 && c.qf_layers > 0 &&
                   c.qf_queries > 0 && c.qf_cross_freq > 0 && c.qf_inter > 0,
               "cgpt_create: bad Q-Former shape");
  CGPT_REQUIRE(c.llm_hidden > 0 && c.llm_heads > 0 && c.llm_hidden % c.llm_heads == 0 && c.llm_layers > 0 &&
                   c.llm_inter > 0 && c.llm_vocab > 0,
               "cgpt_create: bad Llama shape");
  CGPT_REQUIRE(c.max_new_tokens >= 1 && c.n_prefix >= 0 && c.n_suffix >= 0 && c.num_classes >= 2,
               "cgpt_create: bad generation settings");
  cgpt_engine* E = new (std::nothrow) cgpt_engine();
  CGPT_REQUIRE(E != nullptr, "cgpt_create: out of host memory");
  E->c = c;
  E->G = c.img_size / 14;
  E->Pn = E->G * E->G;
  E->T = E->Pn + 1;
  E->vhd = c.vit_dim / c.vit_heads;
  E->qhd = c.qf_hidden / c.qf_heads;
  E->lhd = c.llm_hidden / c.llm_heads;
  E->n_cross = (c.qf_layers + c.qf_cross_freq - 1) / c.qf_cross_freq;
  E->P = c.n_prefix;
  E->ns = c.n_suffix;
  E->Tp = c.qf_queries + c.n_suffix;
  E->Tp_max = E->Tp;
  E->cache_rows = E->P + E->Tp_max + c.max_new_tokens;
  if (cudaStreamCreateWithFlags(&E->cap_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaMallocHost(reinterpret_cast<void**>(&E->h_i32), 3 * 64 * sizeof(int32_t)) != cudaSuccess ||
      cudaMallocHost(reinterpret_cast<void**>(&E->h_f64), 3 * 64 * sizeof(double)) != cudaSuccess) {
    set_last_error("cgpt_create: CUDA stream / pinned host allocation failed (%s)",
                   cudaGetErrorString(cudaGetLastError()));
    cgpt_destroy(E);
    return -2;
  }
  *out = E;
  return 0;
}

int cgpt_destroy(cgpt_handle E) {
  if (E == nullptr) return 0;
  drop_graphs(E);
  if (E->cap_stream) cudaStreamDestroy(E->cap_stream);
  if (E->h_i32) cudaFreeHost(E->h_i32);
  if (E->h_f64) cudaFreeHost(E->h_f64);
  delete E;
  return 0;
}

int cgpt_bind_weight(cgpt_handle E, const char* name, const void* ptr, int64_t rows, int64_t cols, int dtype) {
  CGPT_REQUIRE(E != nullptr && name != nullptr && ptr != nullptr, "cgpt_bind_weight: null argument");
  CGPT_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "cgpt_bind_weight: '%s' is not 16-byte aligned", name);
  Tensor t;
  t.p = ptr; t.rows = rows; t.cols = cols; t.dtype = dtype;
  E->w[name] = t;
  E->resolved = false;
  drop_graphs(E);
  return 0;
}

int cgpt_set_prompt(cgpt_handle E, const int32_t* prefix_ids, const int32_t* suffix_ids) {
  CGPT_REQUIRE(E != nullptr, "cgpt_set_prompt: null handle");
  CGPT_REQUIRE((E->P == 0 || prefix_ids) && (E->ns == 0 || suffix_ids), "cgpt_set_prompt: null ids");
  E->prefix_ids = prefix_ids;
  E->suffix_ids = suffix_ids;
  E->prompt_set = true;
  drop_graphs(E);
  return 0;
}

int cgpt_set_question(cgpt_handle E, const int32_t* suffix_ids, int n_suffix) {
  CGPT_REQUIRE(E != nullptr, "cgpt_set_question: null handle");
  CGPT_REQUIRE(n_suffix >= 0 && n_suffix <= E->c.n_suffix, "cgpt_set_question: %d question tokens, the handle was created for <= %d",
               n_suffix, E->c.n_suffix);
  CGPT_REQUIRE(n_suffix == 0 || suffix_ids != nullptr, "cgpt_set_question: null ids");
  // graphs are keyed by (question length, ids buffer): a caller that rewrites ONE device buffer in place replays the
  // graphs captured for that length, a new buffer or length captures its own set
  E->suffix_ids = suffix_ids;
  E->ns = n_suffix;
  E->Tp = E->c.qf_queries + n_suffix;
  return 0;
}

int cgpt_set_answer_table(cgpt_handle E, const uint64_t* table_keys, const int32_t* table_vals, int capacity) {
  CGPT_REQUIRE(E != nullptr && table_keys && table_vals, "cgpt_set_answer_table: null argument");
  CGPT_REQUIRE(capacity > 0 && (capacity & (capacity - 1)) == 0, "cgpt_set_answer_table: capacity must be a power of 2");
  E->table_keys = table_keys;
  E->table_vals = table_vals;
  E->table_cap = capacity;
  return 0;
}

int cgpt_workspace_bytes(cgpt_handle E, int max_batch, int encoder_only, int64_t* bytes) {
  CGPT_REQUIRE(E != nullptr && bytes != nullptr && max_batch > 0, "cgpt_workspace_bytes: bad argument");
  *bytes = layout(E, max_batch, encoder_only != 0, nullptr, nullptr);
  return 0;
}

int cgpt_bind_workspace(cgpt_handle E, void* workspace, int64_t bytes, int max_batch, int encoder_only,
                        void* stream) {
  CGPT_REQUIRE(E != nullptr && workspace != nullptr && max_batch > 0, "cgpt_bind_workspace: bad argument");
  CGPT_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "cgpt_bind_workspace: workspace must be 256-byte aligned");
  const long long need = layout(E, max_batch, encoder_only != 0, nullptr, nullptr);
  CGPT_REQUIRE(bytes >= need, "cgpt_bind_workspace: %lld bytes given, %lld needed for batch %d",
               static_cast<long long>(bytes), need, max_batch);
  CGPT_TRY(resolve(E));
  drop_graphs(E);
  E->ws = static_cast<char*>(workspace);
  E->ws_bytes = bytes;
  E->ws_B = max_batch;
  E->ws_encoder_only = encoder_only != 0;
  layout(E, max_batch, E->ws_encoder_only, E->ws, &E->b);
  if (!E->ws_encoder_only) {
    CGPT_REQUIRE(E->prompt_set, "cgpt_bind_workspace: prompt ids are not set (cgpt_set_prompt)");
    CGPT_TRY(build_prefix(E, static_cast<cudaStream_t>(stream)));
  }
  return 0;
}

int cgpt_vit_forward(cgpt_handle E, const void* patches, int B, void* out_tokens, void* stream) {
  CGPT_TRY(check_ready(E, B, false));
  CGPT_REQUIRE(patches && out_tokens, "cgpt_vit_forward: null buffer");
  return vit_forward(E, patches, B, out_tokens, static_cast<cudaStream_t>(stream));
}

int cgpt_qformer_forward(cgpt_handle E, const void* tokens, int B, void* out_queries, void* out_llm_embeds,
                         void* stream) {
  CGPT_TRY(check_ready(E, B, false));
  CGPT_REQUIRE(tokens && out_queries, "cgpt_qformer_forward: null buffer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CGPT_TRY(qformer_forward(E, tokens, B, out_queries, s));
  if (out_llm_embeds)   // MiniGPT4.encode_img's inputs_llama (minigpt4.py:141)
    CGPT_TRY(gemm(out_queries, E->c.qf_hidden, E->proj_w, B * E->c.qf_queries, E->c.llm_hidden, E->c.qf_hidden,
                  Epi(out_llm_embeds, E->c.llm_hidden, CGPT_DT_BF16).bias(E->proj_b), s));
  return 0;
}

int cgpt_llm_prefill_decode(cgpt_handle E, const void* queries, int B, int32_t* out_ids, float* out_top2_margin,
                            int* out_steps, void* stream) {
  CGPT_TRY(check_ready(E, B, true));
  CGPT_REQUIRE(queries && out_ids, "cgpt_llm_prefill_decode: null buffer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CGPT_TRY(llm_generate(E, queries, B, s));
  const size_t n = static_cast<size_t>(B) * E->c.max_new_tokens * 4;
  CGPT_CHECK_CUDA(cudaMemcpyAsync(out_ids, E->b.ids, n, cudaMemcpyDeviceToDevice, s));
  if (out_top2_margin) CGPT_CHECK_CUDA(cudaMemcpyAsync(out_top2_margin, E->b.margin, n, cudaMemcpyDeviceToDevice, s));
  if (out_steps) *out_steps = E->last_steps;
  return 0;
}

// ids outside [0, vocab) -> pad id for the embedding lookup: -100 (ignored target), and ids >= vocab, which the loss
// kernel does not score either (ce_rows_kernel) - callers validate user-supplied answers (native.py, train.py)
__global__ void clamp_ids_kernel(const int32_t* __restrict__ src, int32_t* __restrict__ dst, int n, int pad_id,
                                 int vocab) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = (src[i] < 0 || src[i] >= vocab) ? pad_id : src[i];
}

int cgpt_lm_loss(cgpt_handle E, const void* patches, int B, const int32_t* answer_ids, int na, float* out_token_loss,
                 float* out_mean_count, void* stream) {
  CGPT_REQUIRE(E && patches && answer_ids && out_token_loss, "cgpt_lm_loss: null argument");
  const cgpt_model_config& c = E->c;
  CGPT_REQUIRE(na >= 1 && na <= c.max_new_tokens, "cgpt_lm_loss: %d answer tokens, the KV cache holds %d", na, c.max_new_tokens);
  const int Tl = E->Tp + na;                                   // rows per sample: image + suffix + answer
  const long long max_b1 = static_cast<long long>(E->ws_B) * E->Tp / Tl, max_b2 = E->ws_B / na;
  CGPT_TRY(check_ready(E, 1, true));
  CGPT_REQUIRE(B >= 1 && B <= max_b1 && B <= max_b2 && B * na <= E->ws_B,
               "cgpt_lm_loss: batch %d too large for the bound workspace (max %lld)", B, max_b1 < max_b2 ? max_b1 : max_b2);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const Buffers& b = E->b;
  const int nq = c.qf_queries, P = E->P, Hd = c.llm_hidden, ns = E->ns;
  const int M = B * Tl;
  const long long layer_stride = static_cast<long long>(E->ws_B) * E->cache_rows * Hd;
  CGPT_TRY(vit_forward(E, patches, B, b.v_out, s));
  CGPT_TRY(qformer_forward(E, b.v_out, B, b.q_h, s));
  // [image | suffix | answer] rows of every sample
  CGPT_TRY(gemm(b.q_h, c.qf_hidden, E->proj_w, B * nq, Hd, c.qf_hidden,
                Epi(b.l_res, Hd, CGPT_DT_F32).bias(E->proj_b).remap(nq, Tl, 0), s));
  if (ns > 0)
    CGPT_TRY(gather_rows(E->emb, Hd, E->suffix_ids, ns, B * ns, Hd, b.l_res, Hd, CGPT_DT_F32, ns, Tl, nq, c.llm_vocab, s));
  int32_t* ids = b.ids;                                        // [B*na] <= ws_B * max_new_tokens
  clamp_ids_kernel<<<(B * na + 255) / 256, 256, 0, s>>>(answer_ids, ids, B * na, c.pad_id, c.llm_vocab);
  CGPT_CHECK_CUDA(cudaGetLastError());
  count_launch();
  CGPT_TRY(gather_rows(E->emb, Hd, ids, B * na, B * na, Hd, b.l_res, Hd, CGPT_DT_F32, na, Tl, nq + ns, c.llm_vocab, s));
  // the KV cache has P + Tp + max_new rows per sample: rows [P, P + Tl) are (re)written here
  CGPT_TRY(llm_layers(E, M, Tl, B, b.l_res, b.l_xn, b.l_qkv, b.l_att, b.l_act, b.kc, b.vc, layer_stride, P, P,
                      E->cache_rows, 0, s));
  // position (nq + ns - 1 + j) predicts answer token j: final norm + lm_head on those B*na rows only
  CGPT_TRY(norm_rows(b.l_res, Hd, CGPT_DT_F32, E->llm_norm, nullptr, c.llm_rms_eps, B * na, Hd, b.l_xn, Hd, CGPT_DT_BF16, 1,
                     na, Tl, nq + ns - 1, s));
  CGPT_TRY(gemm(b.l_xn, Hd, E->head_w, B * na, c.llm_vocab, Hd, Epi(b.l_logits, c.llm_vocab, CGPT_DT_F32), s));
  CGPT_TRY(ce_loss(static_cast<const float*>(b.l_logits), c.llm_vocab, B * na, c.llm_vocab, answer_ids, out_token_loss,
                   out_mean_count, E->label_smoothing, s));
  return 0;
}

int cgpt_noisy_labels(cgpt_handle E, const float* x, const cgpt_noise_spec* noise, uint64_t first_sample, int B,
                      int32_t* labels, void* stream) {
  CGPT_REQUIRE(E && x && noise && labels, "cgpt_noisy_labels: null argument");
  return noisy_labels(E, x, noise, first_sample, B, labels, static_cast<cudaStream_t>(stream));
}

int cgpt_sample_noise(cgpt_handle E, const float* x, const cgpt_noise_spec* noise, int64_t base, int64_t num,
                      int batch_size, int64_t split, int rank, int world, void* comm, int64_t* counts,
                      int32_t* invalid, void* stream) {
  return sample_noise(E, x, noise, base, num, batch_size, split, rank, world, comm,
                      reinterpret_cast<long long*>(counts), invalid, static_cast<cudaStream_t>(stream));
}

int cgpt_certify(cgpt_handle E, const float* x, const cgpt_noise_spec* noise, int64_t n0, int64_t n, double alpha,
                 int batch_size, int rank, int world, void* comm, int* out_label, double* out_radius,
                 double* out_detail, void* stream) {
  CGPT_REQUIRE(E && out_label && out_radius, "cgpt_certify: null argument");
  CGPT_REQUIRE(n0 > 0 && n > 0, "cgpt_certify: n0 and n must be positive");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // the n0 selection draws and the n estimation draws are independent and keyed by global sample index:
  // ONE sharded pass over [0, n0 + n), two count vectors, one all-reduce (smoothing.py:44,48)
  CGPT_TRY(sample_noise(E, x, noise, 0, n0 + n, batch_size, n0, rank, world, comm, E->b.counts, nullptr, s));
  const int nc = E->c.num_classes;
  const double* lut = (E->lut != nullptr && E->lut_n == n && E->lut_alpha == alpha) ? E->lut : nullptr;
  CGPT_TRY(certify_tail(E->b.counts, E->b.counts + nc, nc, n, alpha, noise->sigma, lut, E->b.tail_label,
                        E->b.tail_stats, s));
  CGPT_CHECK_CUDA(cudaMemcpyAsync(E->h_i32, E->b.tail_label, 2 * 4, cudaMemcpyDeviceToHost, s));
  CGPT_CHECK_CUDA(cudaMemcpyAsync(E->h_f64, E->b.tail_stats, 3 * 8, cudaMemcpyDeviceToHost, s));
  CGPT_CHECK_CUDA(cudaStreamSynchronize(s));   // the one device->host read of the call
  *out_label = E->h_i32[0];
  *out_radius = E->h_i32[0] < 0 ? 0.0 : E->h_f64[0];
  if (out_detail) {
    out_detail[0] = static_cast<double>(E->h_i32[1]);
    out_detail[1] = E->h_f64[1];
    out_detail[2] = E->h_f64[2];
  }
  return 0;
}

// Smooth.certify of K images in shared passes: every pass holds this rank's next draws of ALL K images (K * per rows),
// so the per-rank batch stays large when the draws of an image are sharded over many GPUs (BASELINE.json configs[2]:
// 64 images, N = 1000, 8 GPUs -> 138 draws per rank and image).  Draw i of image k is Philox (seed, stream_id + k, i)
// whatever K, batch size and world size are: per-image counts equal K separate cgpt_certify calls bit for bit.
int cgpt_certify_batch(cgpt_handle E, const float* const* xs, int K, const cgpt_noise_spec* noise, int64_t n0, int64_t n,
                       double alpha, int batch_size, int rank, int world, void* comm, int* out_labels, double* out_radii,
                       double* out_detail, void* stream) {
  CGPT_REQUIRE(E && xs && noise && out_labels && out_radii, "cgpt_certify_batch: null argument");
  CGPT_REQUIRE(n0 > 0 && n > 0 && batch_size > 0 && world > 0 && rank >= 0 && rank < world, "cgpt_certify_batch: bad sizes");
  CGPT_REQUIRE(K >= 1 && K <= E->max_images && K <= 64, "cgpt_certify_batch: %d images, the workspace serves %d (<= 64)", K,
               E->max_images);
  CGPT_TRY(check_ready(E, 1, true));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const Buffers& b = E->b;
  const int nc = E->c.num_classes;
  const long long img_elems = 3LL * E->c.img_size * E->c.img_size;
  CGPT_CHECK_CUDA(cudaMemsetAsync(b.counts, 0, static_cast<size_t>(K) * 2 * nc * 8, s));
  CGPT_CHECK_CUDA(cudaMemsetAsync(b.invalid, 0, 4, s));
  for (int k = 0; k < K; ++k) {
    CGPT_REQUIRE(xs[k] != nullptr, "cgpt_certify_batch: image %d is null", k);
    CGPT_CHECK_CUDA(cudaMemcpyAsync(b.x_static + k * img_elems, xs[k], img_elems * 4, cudaMemcpyDefault, s));
  }
  const long long num = n0 + n;
  const long long lo = (num * rank) / world, hi = (num * (rank + 1)) / world;
  if (batch_size > E->ws_B) batch_size = E->ws_B;
  long long per_max = batch_size / K;
  CGPT_REQUIRE(per_max >= 1, "cgpt_certify_batch: batch_size %d holds no draw of each of the %d images", batch_size, K);
  const long long n_chunks = hi > lo ? (hi - lo + per_max - 1) / per_max : 0;
  long long first = lo;
  for (long long ci = 0; ci < n_chunks; ++ci) {
    const int per = static_cast<int>((hi - first + (n_chunks - ci) - 1) / (n_chunks - ci));   // balanced, each <= per_max
    CGPT_TRY(noisy_labels_images(E, noise, static_cast<uint64_t>(first), K, per, b.labels, s));
    CGPT_TRY(label_hist_images(b.labels, K * per, per, first, n0, nc, b.counts, b.invalid, s));
    first += per;
  }
  if (world > 1 && comm != nullptr)
    CGPT_TRY(cgpt_allreduce_counts(reinterpret_cast<int64_t*>(b.counts), K * 2 * nc, comm, s));
  const double* lut = (E->lut != nullptr && E->lut_n == n && E->lut_alpha == alpha) ? E->lut : nullptr;
  CGPT_TRY(certify_tail_batch(b.counts, K, nc, n, alpha, noise->sigma, lut, b.tail_label, b.tail_stats, s));
  CGPT_CHECK_CUDA(cudaMemcpyAsync(E->h_i32, b.tail_label, K * 3 * 4, cudaMemcpyDeviceToHost, s));
  CGPT_CHECK_CUDA(cudaMemcpyAsync(E->h_f64, b.tail_stats, K * 3 * 8, cudaMemcpyDeviceToHost, s));
  CGPT_CHECK_CUDA(cudaStreamSynchronize(s));   // the one device->host read of the call
  for (int k = 0; k < K; ++k) {
    out_labels[k] = E->h_i32[3 * k];
    out_radii[k] = E->h_i32[3 * k] < 0 ? 0.0 : E->h_f64[3 * k];
    if (out_detail) {
      out_detail[3 * k] = static_cast<double>(E->h_i32[3 * k + 1]);
      out_detail[3 * k + 1] = E->h_f64[3 * k + 1];
      out_detail[3 * k + 2] = E->h_f64[3 * k + 2];
    }
  }
  return 0;
}

int cgpt_set_radius_lut(cgpt_handle E, int64_t n, double alpha, const double* lut) {
  CGPT_REQUIRE(E != nullptr, "cgpt_set_radius_lut: null handle");
  CGPT_REQUIRE(lut == nullptr || (n > 0 && alpha > 0.0 && alpha < 1.0), "cgpt_set_radius_lut: bad (n, alpha)");
  E->lut = lut;
  E->lut_n = n;
  E->lut_alpha = alpha;
  return 0;
}

int cgpt_predict(cgpt_handle E, const float* x, const cgpt_noise_spec* noise, int64_t n, double alpha,
                 int batch_size, int rank, int world, void* comm, int* out_label, double* out_detail,
                 void* stream) {
  CGPT_REQUIRE(E && out_label, "cgpt_predict: null argument");
  CGPT_REQUIRE(n > 0, "cgpt_predict: n must be positive");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CGPT_TRY(sample_noise(E, x, noise, 0, n, batch_size, -1, rank, world, comm, E->b.counts, nullptr, s));
  CGPT_TRY(predict_tail(E->b.counts, E->c.num_classes, alpha, E->b.tail_label, E->b.tail_stats, s));
  CGPT_CHECK_CUDA(cudaMemcpyAsync(E->h_i32, E->b.tail_label, 3 * 4, cudaMemcpyDeviceToHost, s));
  CGPT_CHECK_CUDA(cudaMemcpyAsync(E->h_f64, E->b.tail_stats, 3 * 8, cudaMemcpyDeviceToHost, s));
  CGPT_CHECK_CUDA(cudaStreamSynchronize(s));
  *out_label = E->h_i32[0];
  if (out_detail) out_detail[0] = E->h_f64[0];
  return 0;
}

int cgpt_last_counts(cgpt_handle E, const int64_t** counts) {
  CGPT_REQUIRE(E && counts && E->ws, "cgpt_last_counts: no workspace bound");
  *counts = reinterpret_cast<const int64_t*>(E->b.counts);
  return 0;
}

int cgpt_last_decode_steps(cgpt_handle E) { return E ? E->last_steps : -1; }

int cgpt_set_option(cgpt_handle E, const char* key, int value) {
  CGPT_REQUIRE(E != nullptr && key != nullptr, "cgpt_set_option: null argument");
  if (!strcmp(key, "use_graphs")) E->c.use_graphs = value != 0;
  else if (!strcmp(key, "early_exit")) E->c.early_exit = value != 0;
  else if (!strcmp(key, "label_smoothing_permille")) {
    CGPT_REQUIRE(value >= 0 && value < 1000, "cgpt_set_option: label_smoothing_permille %d outside [0, 1000)", value);
    E->label_smoothing = static_cast<float>(value) / 1000.f;
  }
  else if (!strcmp(key, "max_images")) {
    // images per pass of cgpt_certify_batch; changes the workspace layout: set it BEFORE cgpt_workspace_bytes / bind
    CGPT_REQUIRE(value >= 1 && value <= 64, "cgpt_set_option: max_images %d outside [1, 64]", value);
    if (value != E->max_images) {   // the layout changes: the caller must size and bind the workspace again
      drop_graphs(E);
      E->ws = nullptr;
      E->ws_B = 0;
      E->max_images = value;
    }
  }
  else CGPT_REQUIRE(false, "cgpt_set_option: unknown key '%s'", key);
  return 0;
}

// ---------------------------------------------------------------- NCCL count all-reduce
int cgpt_comm_unique_id(void* id128) {
  CGPT_REQUIRE(id128 != nullptr, "cgpt_comm_unique_id: null argument");
  CGPT_TRY(load_nccl());
  CGPT_CHECK_NCCL(g_nccl.GetUniqueId(id128));
  return 0;
}

int cgpt_comm_init(const void* id128, int rank, int world, void** comm) {
  CGPT_REQUIRE(id128 && comm && world > 0 && rank >= 0 && rank < world, "cgpt_comm_init: bad argument");
  CGPT_TRY(load_nccl());
  Id128 id;
  memcpy(&id, id128, sizeof(id));
  CGPT_CHECK_NCCL(g_nccl.CommInitRank(comm, world, id, rank));
  return 0;
}

int cgpt_comm_destroy(void* comm) {
  if (comm == nullptr) return 0;
  CGPT_TRY(load_nccl());
  CGPT_CHECK_NCCL(g_nccl.CommDestroy(comm));
  return 0;
}

int cgpt_allreduce_f32(float* values, int64_t n, void* comm, void* stream) {
  CGPT_REQUIRE(values && comm && n > 0, "cgpt_allreduce_f32: bad argument");
  CGPT_TRY(load_nccl());
  // ncclFloat32 = 7, ncclSum = 0
  CGPT_CHECK_NCCL(g_nccl.AllReduce(values, values, static_cast<size_t>(n), 7, 0, comm, static_cast<cudaStream_t>(stream)));
  return 0;
}

int cgpt_allreduce_counts(int64_t* counts, int n, void* comm, void* stream) {
  CGPT_REQUIRE(counts && comm && n > 0, "cgpt_allreduce_counts: bad argument");
  CGPT_TRY(load_nccl());
  // ncclInt64 = 4, ncclSum = 0
  CGPT_CHECK_NCCL(g_nccl.AllReduce(counts, counts, static_cast<size_t>(n), 4, 0, comm,
                                   static_cast<cudaStream_t>(stream)));
  return 0;
}

}  // extern "C"
