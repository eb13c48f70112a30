/* Plain C99 client of libcgpt.so: what a cgo / JNI / Rust-FFI binding of the reference would link against.
 * No Python, no torch.  On a host without a GPU every compute entry point must FAIL with a message (there is no CPU
 * fallback); the host-side helpers (ABI version, answer hash) work anywhere.  Exit code 0 = all checks passed.
 * Built and run by tests/test_abi_cpu.py::test_plain_c_client_links_and_fails_loudly_without_a_gpu. */
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "cgpt.h"

#define CHECK(cond)                                                        \
  do {                                                                     \
    if (!(cond)) {                                                         \
      fprintf(stderr, "abi_client: check failed at line %d: %s\n", __LINE__, #cond); \
      return 1;                                                            \
    }                                                                      \
  } while (0)

int main(int argc, char** argv) {
  int expect_gpu = argc > 1 && strcmp(argv[1], "--gpu") == 0;
  CHECK(cgpt_abi_version() == CGPT_ABI_VERSION);
  CHECK(cgpt_launch_count() == 0);

  /* answer hash: stops at EOS, drops ids 0/1/2 (pad, bos, eos) - minigpt_base.py:438-446 */
  int32_t a[5] = {1, 3869, 2, 77, 0};
  int32_t b[1] = {3869};
  int32_t c[2] = {3869, 77};
  CHECK(cgpt_answer_hash(a, 5, 2) == cgpt_answer_hash(b, 1, 2));
  CHECK(cgpt_answer_hash(a, 5, 2) != cgpt_answer_hash(c, 2, 2));

  /* argument validation happens before any device work and reports through cgpt_last_error() */
  cgpt_model_config cfg;
  memset(&cfg, 0, sizeof(cfg));
  cgpt_handle h = NULL;
  CHECK(cgpt_create(NULL, &h) != 0 && strlen(cgpt_last_error()) > 0);
  CHECK(cgpt_create(&cfg, &h) != 0 && h == NULL);            /* img_size 0 */
  CHECK(strstr(cgpt_last_error(), "img_size") != NULL);
  CHECK(cgpt_gemm_bf16(NULL, 0, NULL, 0, 0, 0, 0, NULL, 0, NULL) != 0);
  CHECK(cgpt_attention(NULL, NULL) != 0);
  CHECK(cgpt_destroy(NULL) == 0);

  /* a well-formed tiny configuration: needs a CUDA context */
  cfg.img_size = 28; cfg.vit_dim = 64; cfg.vit_depth = 2; cfg.vit_heads = 2; cfg.vit_mlp = 128;
  cfg.vit_eps = 1e-6f; cfg.ln_vision_eps = 1e-5f;
  cfg.qf_hidden = 64; cfg.qf_layers = 2; cfg.qf_heads = 2; cfg.qf_inter = 128; cfg.qf_queries = 8; cfg.qf_cross_freq = 2;
  cfg.qf_eps = 1e-12f;
  cfg.llm_hidden = 128; cfg.llm_layers = 2; cfg.llm_heads = 2; cfg.llm_inter = 256; cfg.llm_vocab = 128; cfg.llm_rms_eps = 1e-5f;
  cfg.eos_id = 2; cfg.pad_id = 0; cfg.n_prefix = 3; cfg.n_suffix = 4; cfg.max_new_tokens = 2; cfg.min_length = 1;
  cfg.num_classes = 6; cfg.early_exit = 1; cfg.use_graphs = 1;
  int rc = cgpt_create(&cfg, &h);
  if (expect_gpu) {
    CHECK(rc == 0 && h != NULL);
    CHECK(cgpt_last_decode_steps(h) <= 0);
    CHECK(cgpt_set_option(h, "no_such_option", 1) != 0 && strstr(cgpt_last_error(), "no_such_option") != NULL);
    CHECK(cgpt_destroy(h) == 0);
  } else {
    CHECK(rc != 0 && h == NULL && strlen(cgpt_last_error()) > 0);
    float x[4] = {0.f, 0.f, 0.f, 0.f};
    long long counts[6];
    cgpt_noise_spec ns;
    memset(&ns, 0, sizeof(ns));
    /* without a handle nothing computes */
    CHECK(cgpt_sample_noise(NULL, x, &ns, 0, 8, 4, -1, 0, 1, NULL, (int64_t*)counts, NULL, NULL) != 0);
  }
  printf("abi_client ok (%s)\n", expect_gpu ? "gpu" : "cpu-only host");
  return 0;
}
