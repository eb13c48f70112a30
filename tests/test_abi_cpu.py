"""CPU tests: libcgpt.so builds/loads and exports every symbol include/cgpt.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "cgpt.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cgpt_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_hot_path_entry_points():
    names = _declared_symbols()
    for must in ["cgpt_noise_patchify", "cgpt_gemm_bf16", "cgpt_label_hist", "cgpt_certify_tail",
                 "cgpt_predict_tail", "cgpt_answer_labels", "cgpt_last_error",
                 # the per-subsystem entry points SURVEY.md 8(b) lists
                 "cgpt_create", "cgpt_destroy", "cgpt_bind_weight", "cgpt_workspace_bytes", "cgpt_vit_forward",
                 "cgpt_qformer_forward", "cgpt_llm_prefill_decode", "cgpt_sample_noise", "cgpt_certify",
                 "cgpt_predict", "cgpt_allreduce_counts"]:
        assert must in names


def test_library_exports_every_declared_symbol():
    from certifiedgpt_b200 import build
    lib_path = build.build()  # nvcc cross-compiles without a GPU; no-op when current
    lib = ctypes.CDLL(lib_path)
    missing = [n for n in _declared_symbols() if not hasattr(lib, n)]
    assert not missing, f"libcgpt.so lacks {missing}"
    lib.cgpt_abi_version.restype = ctypes.c_int
    assert lib.cgpt_abi_version() == 3


def test_no_cpu_fallback_for_compute():
    """Without a CUDA tensor the product path must raise, not fall back."""
    import torch
    from certifiedgpt_b200 import _lib as L
    from certifiedgpt_b200.randomized_smoothing.smoothing import Smooth
    s = Smooth(torch.nn.Identity(), 3, 0.25)
    with pytest.raises(L.CgptError):
        s._sample_noise(torch.zeros(3, 8, 8), 4, 4)


def test_answer_hash_canonicalisation_host():
    from certifiedgpt_b200 import _lib as L
    # stop at EOS (2), drop special ids 0/1/2 (decode(skip_special_tokens=True), minigpt_base.py:442)
    assert L.answer_hash([5, 6, 2, 9]) == L.answer_hash([0, 5, 1, 6]) == L.answer_hash([5, 6])
    assert L.answer_hash([5, 6]) != L.answer_hash([6, 5])
    assert L.answer_hash([5]) != L.answer_hash([5, 5])


def _struct_fields(struct):
    """field names of a C struct in include/cgpt.h, in declaration order"""
    src = open(os.path.join(ROOT, "include", "cgpt.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (struct, struct), src, flags=re.S).group(1)
    out = []
    for decl in body.split(";"):
        for declarator in decl.split(","):
            names = re.findall(r"[A-Za-z_]\w*", re.sub(r"\[.*?\]", "", declarator))
            if names:
                out.append(names[-1])
    return out


def test_ctypes_mirrors_match_the_header_structs():
    """ctypes mirrors of every struct that crosses the ABI: one field per header field, same order."""
    from certifiedgpt_b200 import _lib as L
    from certifiedgpt_b200 import native as N
    for struct, mirror in [("cgpt_model_config", N.ModelConfigC), ("cgpt_noise_spec", N.NoiseSpecC),
                           ("cgpt_gemm_epilogue", L.GemmEpilogue), ("cgpt_gemm_rope", L.GemmRope),
                           ("cgpt_attn_args", L.AttnArgs)]:
        assert _struct_fields(struct) == [f[0] for f in mirror._fields_], struct
    assert ctypes.sizeof(N.ModelConfigC) == 4 * len(N.ModelConfigC._fields_)
    assert ctypes.sizeof(N.NoiseSpecC) == 56
    assert ctypes.sizeof(L.GemmRope) == 64 and ctypes.sizeof(L.GemmEpilogue) == 120


def test_native_engine_refuses_to_start_without_a_gpu():
    """cgpt_create needs a CUDA context (stream + pinned scratch): on a CPU-only host it must fail with a
    message, never hand out a handle that would compute elsewhere."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    from certifiedgpt_b200 import native as N
    from certifiedgpt_b200.config import ModelConfig
    h = N.lib()
    c = N.config_struct(ModelConfig.tiny(), 3, 4, 2, 1, 6, True, True)
    hd = ctypes.c_void_p()
    rc = h.cgpt_create(ctypes.byref(c), ctypes.byref(hd))
    assert rc != 0 and not hd.value and len(h.cgpt_last_error()) > 0
    bad = N.config_struct(ModelConfig.tiny(), 3, 4, 0, 1, 6, True, True)    # max_new_tokens = 0
    assert h.cgpt_create(ctypes.byref(bad), ctypes.byref(hd)) == -1


def test_header_is_plain_c99_and_struct_sizes_match_ctypes(tmp_path):
    """include/cgpt.h must be consumable by a C FFI (cgo / JNI / ctypes generators): compile it as C99 with gcc and
    compare sizeof() of every struct that crosses the ABI with the ctypes mirrors."""
    import shutil
    import subprocess
    from certifiedgpt_b200 import _lib as L
    from certifiedgpt_b200 import native as N
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include "cgpt.h"\nint main(void) { printf("%zu %zu %zu %zu %zu\\n", '
                   "sizeof(cgpt_model_config), sizeof(cgpt_noise_spec), sizeof(cgpt_gemm_epilogue), sizeof(cgpt_gemm_rope), "
                   "sizeof(cgpt_attn_args)); return 0; }\n")
    exe = tmp_path / "sizes"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                           str(src), "-o", str(exe)])
    got = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    want = [ctypes.sizeof(t) for t in (N.ModelConfigC, N.NoiseSpecC, L.GemmEpilogue, L.GemmRope, L.AttnArgs)]
    assert got == want


def test_lr_schedule_matches_the_reference_scheduler():
    """LinearWarmupCosineLRScheduler.step (graphs/models/minigpt4/common/optims.py:11-73) on the shipped settings."""
    import math
    from certifiedgpt_b200.agents.minigpt4_finetune_agent import linear_warmup_cosine_lr
    kw = dict(max_epoch=4, iters_per_epoch=53, min_lr=1e-6, init_lr=1e-5, warmup_steps=53, warmup_start_lr=1e-6, warmup_max_lr=1e-5)
    assert linear_warmup_cosine_lr(0, 0, **kw) == pytest.approx(1e-6)
    assert linear_warmup_cosine_lr(0, 26, **kw) == pytest.approx(1e-6 + 9e-6 * 26 / 53)
    assert linear_warmup_cosine_lr(1, 0, **kw) == pytest.approx((1e-5 - 1e-6) * 0.5 * (1 + math.cos(math.pi * 53 / 212)) + 1e-6)
    assert linear_warmup_cosine_lr(3, 52, **kw) < 1.1e-6
    # every (epoch, step) of three settings against a run of the reference's own scheduler (tests/golden/make_ref_lr_fixture.py)
    import json
    for rec in json.load(open(os.path.join(ROOT, "tests", "golden", "ref_lr.json"))):
        kw = rec["settings"]
        got = [linear_warmup_cosine_lr(e, s, **kw) for e in range(kw["max_epoch"]) for s in range(kw["iters_per_epoch"])]
        assert got == pytest.approx(rec["lr"], rel=1e-15, abs=0)


def test_integration_doc_names_every_entry_point():
    """INTEGRATION.md maps each C symbol of include/cgpt.h to the reference lines it replaces."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "cgpt.h")).read()
    doc = open(os.path.join(root, "INTEGRATION.md")).read()
    syms = sorted(set(re.findall(r"\b(cgpt_[a-z0-9_]+)\s*\(", header)))
    assert len(syms) >= 50
    assert [s for s in syms if s not in doc] == []


def test_plain_c_client_links_and_fails_loudly_without_a_gpu(tmp_path):
    """tests/c/abi_client.c: a C99 program linked against libcgpt.so alone (what a cgo / JNI binding links): host helpers
    work, argument errors come back as status + message, and on this CPU-only host cgpt_create refuses to start."""
    import shutil
    import subprocess
    import torch
    from certifiedgpt_b200 import _lib as L
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    libdir = os.path.dirname(L.LIB_PATH)
    exe = tmp_path / "abi_client"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c", "abi_client.c"), "-o", str(exe), "-L", libdir, "-lcgpt",
                           f"-Wl,-rpath,{libdir}"])
    args = [str(exe)] + (["--gpu"] if torch.cuda.is_available() else [])
    r = subprocess.run(args, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert "abi_client ok" in r.stdout
    # the library's only dynamic dependencies are the C/C++ runtimes (CUDA runtime linked statically, NCCL dlopen'ed)
    deps = subprocess.check_output(["ldd", L.LIB_PATH], text=True)
    assert "libtorch" not in deps and "libpython" not in deps and "libnccl" not in deps
