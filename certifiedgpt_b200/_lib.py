"""ctypes binding of libcgpt.so (the C-ABI declared in include/cgpt.h).

There is no CPU fallback: if the library is missing or a call fails, this module raises.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libcgpt.so")

DT_BF16, DT_F32 = 0, 1
ACT_NONE, ACT_GELU, ACT_SWIGLU = 0, 1, 2
NOISE_GAUSSIAN, NOISE_UNIFORM = 0, 1
SPACE_NORMALIZED, SPACE_PIXEL = 0, 1


class CgptError(RuntimeError):
    pass


class GemmEpilogue(C.Structure):
    _fields_ = [
        ("out", C.c_void_p), ("ldo", C.c_int64), ("out_dtype", C.c_int),
        ("bias", C.c_void_p),
        ("resid", C.c_void_p), ("ldr", C.c_int64), ("resid_dtype", C.c_int),
        ("act", C.c_int),
        ("row_add", C.c_void_p), ("ld_row_add", C.c_int64),
        ("row_period", C.c_int), ("row_add_offset", C.c_int),
        ("remap_stride", C.c_int), ("remap_offset", C.c_int),
        ("max_ctas", C.c_int),
    ]


_lib = None


def load(build_if_missing=False):
    """Load libcgpt.so; raises CgptError when it is absent (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if build_if_missing:
            from . import build as _b
            _b.build()
        else:
            raise CgptError(
                f"{LIB_PATH} not found: build it with `python -m certifiedgpt_b200.build` "
                "(there is no CPU fallback for the smoothing hot path)")
    lib = C.CDLL(LIB_PATH)
    lib.cgpt_last_error.restype = C.c_char_p
    lib.cgpt_launch_count.restype = C.c_longlong
    _declare(lib)
    _lib = lib
    return lib


def _declare(lib):
    vp, i64, i32, f32, f64 = C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_double
    u64 = C.c_uint64
    sigs = {
        "cgpt_abi_version": [],
        "cgpt_gemm_bf16": [vp, i64, vp, i64, i32, i32, i32, C.POINTER(GemmEpilogue), i32, vp],
    }
    sigs.update(_EXTRA_SIGS(vp, i64, i32, f32, f64, u64))
    for name, args in sigs.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_int


def _EXTRA_SIGS(vp, i64, i32, f32, f64, u64):
    return {}


def check(rc):
    if rc != 0:
        raise CgptError(f"libcgpt error {rc}: {load().cgpt_last_error().decode()}")


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """data_ptr of a CUDA tensor (or None)"""
    if t is None:
        return None
    assert t.is_cuda, "libcgpt only takes device buffers"
    return C.c_void_p(t.data_ptr())


def launch_count():
    return int(load().cgpt_launch_count())


# ------------------------------------------------------------------------------- GEMM
def gemm(a, w, *, out=None, bias=None, resid=None, act=ACT_NONE, out_dtype=torch.bfloat16,
         row_add=None, row_period=0, row_add_offset=0, remap_stride=0, remap_offset=0,
         out_rows=None, force_bn=0, max_ctas=0):
    """out = epilogue(a @ w.T); a [M,K] bf16 (row stride may exceed K), w [N,K] bf16."""
    lib = load()
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16
    assert a.stride(1) == 1 and w.stride(1) == 1
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K
    n_out = N // 2 if act == ACT_SWIGLU else N
    if out is None:
        rows = out_rows if out_rows is not None else M
        out = torch.empty(rows, n_out, dtype=out_dtype, device=a.device)
    e = GemmEpilogue()
    e.out = out.data_ptr(); e.ldo = out.stride(0)
    e.out_dtype = DT_F32 if out.dtype == torch.float32 else DT_BF16
    e.bias = bias.data_ptr() if bias is not None else None
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == N
    if resid is not None:
        e.resid = resid.data_ptr(); e.ldr = resid.stride(0)
        e.resid_dtype = DT_F32 if resid.dtype == torch.float32 else DT_BF16
    e.act = act
    if row_add is not None:
        assert row_add.dtype == torch.float32
        e.row_add = row_add.data_ptr(); e.ld_row_add = row_add.stride(0)
    e.row_period = row_period; e.row_add_offset = row_add_offset
    e.remap_stride = remap_stride; e.remap_offset = remap_offset
    e.max_ctas = max_ctas
    check(lib.cgpt_gemm_bf16(ptr(a), a.stride(0), ptr(w), w.stride(0), M, N, K,
                             C.byref(e), force_bn, stream_ptr()))
    return out
