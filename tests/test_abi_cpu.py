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
                 "cgpt_predict_tail", "cgpt_answer_labels", "cgpt_last_error"]:
        assert must in names


def test_library_exports_every_declared_symbol():
    from certifiedgpt_b200 import build
    lib_path = build.build()  # nvcc cross-compiles without a GPU; no-op when current
    lib = ctypes.CDLL(lib_path)
    missing = [n for n in _declared_symbols() if not hasattr(lib, n)]
    assert not missing, f"libcgpt.so lacks {missing}"
    lib.cgpt_abi_version.restype = ctypes.c_int
    assert lib.cgpt_abi_version() == 1


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
