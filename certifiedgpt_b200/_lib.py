"""ctypes binding of libcgpt.so (the C-ABI declared in include/cgpt.h).

There is no CPU fallback: if the library is missing or a call fails, this module raises.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# CGPT_LIB=<path>: load another build of the same ABI (A/B runs of two kernel versions inside one gpurun call)
LIB_PATH = os.environ.get("CGPT_LIB") or os.path.join(_HERE, "lib", "libcgpt.so")

DT_BF16, DT_F32 = 0, 1
ACT_NONE, ACT_GELU, ACT_SWIGLU, ACT_QUICKGELU = 0, 1, 2, 3
NOISE_GAUSSIAN, NOISE_UNIFORM = 0, 1
SPACE_NORMALIZED, SPACE_PIXEL = 0, 1


class CgptError(RuntimeError):
    pass


class GemmRope(C.Structure):
    _fields_ = [
        ("T", C.c_int), ("heads", C.c_int), ("pos0", C.c_int),
        ("cos_table", C.c_void_p), ("sin_table", C.c_void_p), ("kcache", C.c_void_p), ("vcache", C.c_void_p),
        ("ld_cache", C.c_int64), ("cache_rows_per_batch", C.c_int), ("cache_row0", C.c_int),
    ]


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
        ("rope", C.POINTER(GemmRope)),
        ("hm_T", C.c_int), ("hm_heads", C.c_int), ("hm_hd", C.c_int),
    ]


class AttnArgs(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("ldq", C.c_int64), ("q_rows_per_batch", C.c_int),
        ("k", C.c_void_p), ("v", C.c_void_p), ("ldk", C.c_int64), ("ldv", C.c_int64),
        ("kv_rows_per_batch", C.c_int),
        ("kp", C.c_void_p), ("vp", C.c_void_p), ("ldkp", C.c_int64), ("ldvp", C.c_int64), ("P", C.c_int),
        ("o", C.c_void_p), ("ldo", C.c_int64),
        ("B", C.c_int), ("H", C.c_int), ("Tq", C.c_int), ("Tk", C.c_int), ("head_dim", C.c_int),
        ("scale", C.c_float), ("causal", C.c_int), ("decode_kernel", C.c_int), ("head_major", C.c_int),
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
    lib.cgpt_answer_hash.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.cgpt_answer_hash.restype = C.c_uint64
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
    u32 = C.c_uint32
    return {
        "cgpt_noise_patchify": [vp, vp, u64, u32, u64, i32, f32, vp, vp, i32, i32, i32, vp, i64, vp],
        "cgpt_noise_patchify_dyn": [vp, vp, i32, vp, vp, i32, i32, i32, vp, i64, vp],
        "cgpt_noise_image": [vp, vp, u64, u32, u64, i32, f32, vp, vp, i32, i32, i32, i32, i32, vp, vp],
        "cgpt_answer_labels": [vp, i32, i32, i32, i32, vp, vp, i32, i32, vp, vp],
        "cgpt_argmax_rows": [vp, i32, i32, i64, i32, vp, vp, vp],
        "cgpt_label_hist": [vp, i32, i32, vp, vp, vp],
        "cgpt_certify_tail": [vp, vp, i32, i64, f64, f64, vp, vp, vp],
        "cgpt_certify_tail_lut": [vp, vp, i32, i64, f64, f64, vp, vp, vp, vp],
        "cgpt_predict_tail": [vp, i32, f64, vp, vp, vp],
        "cgpt_greedy_step": [vp, i32, vp, vp, i32, i32, i32, i32, vp, vp],
        "cgpt_norm_rows": [vp, i64, i32, vp, vp, f32, i32, i32, vp, i64, i32, i32, i32, i32, i32, vp],
        "cgpt_attention": [C.POINTER(AttnArgs), vp],
        "cgpt_rope_split": [vp, i64, i32, i32, i32, i32, i32, vp, vp, vp, vp, i64, i32, i32, vp],
        "cgpt_gather_rows": [vp, i64, vp, i32, i32, i32, vp, i64, i32, i32, i32, i32, i32, vp],
        "cgpt_cosine_rows": [vp, i64, i32, i32, vp, vp, vp],
        "cgpt_ce_loss": [vp, i64, i32, i32, vp, vp, vp, vp],
        "cgpt_ce_loss_smooth": [vp, i64, i32, i32, vp, vp, vp, f32, vp],
        "cgpt_swiglu_fwd": [vp, vp, i64, i32, vp],
        "cgpt_swiglu_bwd": [vp, vp, vp, i64, i32, vp],
        "cgpt_rmsnorm_bwd": [vp, i64, vp, vp, i64, f32, i32, i32, vp, i64, i32, i32, i32, vp],
        "cgpt_rope_bwd_cast": [vp, vp, i32, i32, i32, i32, i32, vp, vp, vp],
        "cgpt_attention_bwd": [vp, i64, vp, vp, i64, i32, vp, i64, vp, i64, vp, i32, i32, i32, i32, i32, f32, vp],
        "cgpt_ce_grad": [vp, i64, i32, i32, vp, vp, vp, i64, vp],
        "cgpt_ce_grad_smooth": [vp, i64, i32, i32, vp, vp, vp, i64, f32, vp],
        "cgpt_cast_rows_f32_bf16": [vp, i64, vp, i64, i32, i32, i32, i32, i32, vp],
        "cgpt_transpose_bf16": [vp, i64, vp, i64, i32, i32, vp],
        "cgpt_colsum_bf16": [vp, i64, i32, i32, vp, vp],
        "cgpt_adamw_step": [vp, vp, vp, vp, vp, i64, f32, f32, f32, f32, f32, i32, f32, vp],
        "cgpt_gemm_profile_begin": [],
        "cgpt_gemm_profile_end": [vp, vp, i32, C.POINTER(C.c_int)],
    }


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
def gemm_profile_start():
    """Bracket every eager GEMM launch (cgpt_gemm_bf16 and the native engine's) with a CUDA-event pair on its
    stream, inside libcgpt (bench.py roofline measurement)."""
    check(load().cgpt_gemm_profile_begin())


def gemm_profile_stop():
    """-> list of (ms, flops, (M, N, K)); synchronises."""
    import numpy as np
    cap = 1 << 16
    ms = np.zeros(cap, dtype=np.float32)
    mnk = np.zeros(3 * cap, dtype=np.int32)
    n = C.c_int(0)
    check(load().cgpt_gemm_profile_end(ms.ctypes.data_as(C.c_void_p), mnk.ctypes.data_as(C.c_void_p), cap,
                                       C.byref(n)))
    k = min(n.value, cap)
    return [(float(ms[i]), 2.0 * float(mnk[3 * i]) * float(mnk[3 * i + 1]) * float(mnk[3 * i + 2]),
             (int(mnk[3 * i]), int(mnk[3 * i + 1]), int(mnk[3 * i + 2]))) for i in range(k)]


def gemm(a, w, *, out=None, bias=None, resid=None, act=ACT_NONE, out_dtype=torch.bfloat16,
         row_add=None, row_period=0, row_add_offset=0, remap_stride=0, remap_offset=0,
         out_rows=None, force_bn=0, max_ctas=0, rope=None, headmajor=None):
    """out = epilogue(a @ w.T); a [M,K] bf16 (row stride may exceed K), w [N,K] bf16.
    headmajor=(T, heads, hd): fused q|k|v projection scattered as [3][M/T][heads][T][hd] into `out` (same bytes)."""
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
    if rope is not None:
        # fused rotary + KV-cache append (128-wide heads): dict(T, heads, pos0, cos, sin, kcache, vcache, cache_rows, cache_row0)
        r = GemmRope()
        r.T, r.heads, r.pos0 = rope["T"], rope["heads"], rope["pos0"]
        r.cos_table, r.sin_table = rope["cos"].data_ptr(), rope["sin"].data_ptr()
        r.kcache, r.vcache = rope["kcache"].data_ptr(), rope["vcache"].data_ptr()
        r.ld_cache = rope["kcache"].stride(-2)
        r.cache_rows_per_batch, r.cache_row0 = rope["cache_rows"], rope["cache_row0"]
        e.rope = C.pointer(r)
    if headmajor is not None:
        e.hm_T, e.hm_heads, e.hm_hd = (int(v) for v in headmajor)
        assert out.dtype == torch.bfloat16 and out.is_contiguous() and out.numel() >= M * N
    check(lib.cgpt_gemm_bf16(ptr(a), a.stride(0), ptr(w), w.stride(0), M, N, K,
                             C.byref(e), force_bn, stream_ptr()))
    return out


# ------------------------------------------------------------------------------- noise
BLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)   # processors/base_processor.py:17
BLIP_STD = (0.26862954, 0.26130258, 0.27577711)   # processors/base_processor.py:19


def _f3(vals):
    return (C.c_float * 4)(*[float(v) for v in vals], *([0.0] * (4 - len(vals))))


def noise_patchify(x, B, sigma, *, eps=None, seed=0, stream_id=0, first_sample=0,
                   noise_space=SPACE_NORMALIZED, noise_kind=NOISE_GAUSSIAN,
                   mean=BLIP_MEAN, std=BLIP_STD, out=None):
    """K1: x [3,S,S] fp32 -> bf16 patch rows [B*(S/14)^2, 592]."""
    lib = load()
    assert x.dtype == torch.float32 and x.dim() == 3 and x.shape[0] == 3 and x.is_contiguous()
    S = x.shape[-1]
    assert x.shape[1] == S
    G = S // 14
    if eps is not None:
        assert eps.dtype == torch.float32 and eps.is_contiguous() and tuple(eps.shape) == (B, 3, S, S)
    if out is None:
        out = torch.empty(B * G * G, 592, dtype=torch.bfloat16, device=x.device)
    check(lib.cgpt_noise_patchify(ptr(x), ptr(eps), seed, stream_id, first_sample, B, float(sigma),
                                  _f3(mean), _f3(std), noise_space, noise_kind, S, ptr(out),
                                  out.stride(0), stream_ptr()))
    return out


def noise_patchify_dyn(x, dyn_params, B, *, noise_space=SPACE_NORMALIZED, noise_kind=NOISE_GAUSSIAN,
                       mean=BLIP_MEAN, std=BLIP_STD, out=None):
    """K1 with (seed, first_sample, stream_id, sigma) read from a 24-byte device struct (graph replay)."""
    lib = load()
    S = x.shape[-1]
    check(lib.cgpt_noise_patchify_dyn(ptr(x), ptr(dyn_params), B, _f3(mean), _f3(std), noise_space,
                                      noise_kind, S, ptr(out), out.stride(0), stream_ptr()))
    return out


def pack_noise_dyn(seed, first_sample, stream_id, sigma):
    """Host bytes of the NoiseDyn struct (little endian: u64 seed, u64 first_sample, u32 stream_id, f32 sigma)."""
    import struct
    return struct.pack("<QQIf", int(seed) & 0xFFFFFFFFFFFFFFFF, int(first_sample), int(stream_id) & 0xFFFFFFFF,
                       float(sigma))


def noise_image(x, B, sigma, *, eps=None, seed=0, stream_id=0, first_sample=0,
                noise_space=SPACE_NORMALIZED, noise_kind=NOISE_GAUSSIAN, mean=None, std=None,
                out=None):
    """Generic path: x [C,H,W] fp32 -> [B,C,H,W] fp32 = x + sigma*eps (optionally normalised)."""
    lib = load()
    assert x.dtype == torch.float32 and x.dim() == 3 and x.is_contiguous()
    Cc, H, W = x.shape
    if eps is not None:
        assert eps.dtype == torch.float32 and eps.is_contiguous() and tuple(eps.shape) == (B, Cc, H, W)
    if out is None:
        out = torch.empty(B, Cc, H, W, dtype=torch.float32, device=x.device)
    m = _f3(mean) if mean is not None else None
    sd = _f3(std) if std is not None else None
    check(lib.cgpt_noise_image(ptr(x), ptr(eps), seed, stream_id, first_sample, B, float(sigma),
                               m, sd, noise_space, noise_kind, Cc, H, W, ptr(out), stream_ptr()))
    return out


# ------------------------------------------------------------------------------- labels / stats
def answer_hash(ids, eos_id=2):
    arr = (C.c_int32 * len(ids))(*[int(i) for i in ids])
    return int(load().cgpt_answer_hash(arr, len(ids), eos_id))


def build_answer_table(entries, eos_id=2, device="cuda"):
    """entries: iterable of (token_id_sequence, label).  Returns (keys u64-as-i64, vals i32)."""
    entries = list(entries)
    cap = 16
    while cap < 2 * max(1, len(entries)):
        cap *= 2
    keys = [0] * cap
    vals = [-1] * cap
    for seq, label in entries:
        h = answer_hash(seq, eos_id)
        slot = (h ^ (h >> 32)) & 0xffffffff & (cap - 1)
        while keys[slot] != 0 and keys[slot] != h:
            slot = (slot + 1) & (cap - 1)
        if keys[slot] == h and vals[slot] != label:
            raise ValueError(f"answer table: two labels for one canonical answer {seq}")
        keys[slot] = h
        vals[slot] = int(label)
    import numpy as np
    k = torch.from_numpy(np.array(keys, dtype=np.uint64).view(np.int64)).to(device)
    v = torch.tensor(vals, dtype=torch.int32, device=device)
    return k, v


def answer_labels(ids, table_keys, table_vals, other_label, eos_id=2, out=None):
    lib = load()
    assert ids.dtype == torch.int32 and ids.dim() == 2 and ids.stride(1) == 1
    B, max_new = ids.shape
    if out is None:
        out = torch.empty(B, dtype=torch.int32, device=ids.device)
    check(lib.cgpt_answer_labels(ptr(ids), B, max_new, ids.stride(0), eos_id, ptr(table_keys),
                                 ptr(table_vals), table_keys.numel(), other_label, ptr(out),
                                 stream_ptr()))
    return out


def argmax_rows(logits, suppress_col=-1, want_margin=False):
    lib = load()
    assert logits.dtype == torch.float32 and logits.dim() == 2 and logits.stride(1) == 1
    rows, cols = logits.shape
    idx = torch.empty(rows, dtype=torch.int32, device=logits.device)
    margin = torch.empty(rows, dtype=torch.float32, device=logits.device) if want_margin else None
    check(lib.cgpt_argmax_rows(ptr(logits), rows, cols, logits.stride(0), suppress_col, ptr(idx),
                               ptr(margin), stream_ptr()))
    return (idx, margin) if want_margin else idx


def label_hist(labels, counts, invalid=None):
    """counts (int64 [num_classes], device) += histogram(labels); asynchronous."""
    lib = load()
    assert labels.dtype == torch.int32 and counts.dtype == torch.int64 and labels.is_contiguous()
    check(lib.cgpt_label_hist(ptr(labels), labels.numel(), counts.numel(), ptr(counts), ptr(invalid),
                              stream_ptr()))
    return counts


_RADIUS_LUTS = {}


def radius_lut_host(n, alpha):
    """float64 [2*(n+1)]: pABar(nA) = Smooth._lower_confidence_bound(nA, n, alpha) and norm.ppf(pABar(nA)), nA = 0..n,
    by the SciPy calls the reference makes once per image on the host (smoothing.py:55,117:
    proportion_confint(NA, N, alpha=2*alpha, method="beta")[0] is beta.ppf(alpha, NA, N-NA+1) with 0 at NA = 0;
    radius = sigma * norm.ppf(pABar)).  Tabulated once per (n, alpha); the Monte-Carlo loop itself never touches the
    host.  Entries with pABar < 0.5 keep ppf = 0 (the reference abstains there and never evaluates it)."""
    import numpy as np
    from scipy.stats import beta, norm
    n = int(n)
    na = np.arange(1, n + 1, dtype=np.int64)
    pab = np.zeros(n + 1, dtype=np.float64)
    pab[1:] = beta.ppf(float(alpha), na, n - na + 1)
    ppf = np.zeros(n + 1, dtype=np.float64)
    ok = pab >= 0.5
    ppf[ok] = norm.ppf(pab[ok])
    return np.concatenate([pab, ppf])


def radius_lut(n, alpha, device):
    """Device copy of `radius_lut_host` (cached per (n, alpha, device))."""
    dev = torch.device(device)
    key = (int(n), float(alpha), dev.type, dev.index if dev.index is not None else torch.cuda.current_device())
    t = _RADIUS_LUTS.get(key)
    if t is None:
        t = torch.from_numpy(radius_lut_host(n, alpha)).to(dev)
        _RADIUS_LUTS[key] = t
    return t


def certify_tail(counts_sel, counts_est, n, alpha, sigma, lut=None):
    """Device tail of Smooth.certify; returns (label_dev int32[2], stats_dev f64[3]).  lut: `radius_lut(n, alpha)`
    for a pABar / radius bit-identical to the reference's SciPy tail; None = device bisection + AS241."""
    lib = load()
    assert counts_sel.dtype == torch.int64 and counts_est.dtype == torch.int64
    lab = torch.empty(2, dtype=torch.int32, device=counts_sel.device)
    st = torch.empty(3, dtype=torch.float64, device=counts_sel.device)
    if lut is not None:
        assert lut.dtype == torch.float64 and lut.numel() == 2 * (int(n) + 1) and lut.is_cuda
        check(lib.cgpt_certify_tail_lut(ptr(counts_sel), ptr(counts_est), counts_sel.numel(), int(n),
                                        float(alpha), float(sigma), ptr(lut), ptr(lab), ptr(st), stream_ptr()))
    else:
        check(lib.cgpt_certify_tail(ptr(counts_sel), ptr(counts_est), counts_sel.numel(), int(n),
                                    float(alpha), float(sigma), ptr(lab), ptr(st), stream_ptr()))
    return lab, st


def predict_tail(counts, alpha):
    lib = load()
    assert counts.dtype == torch.int64
    lab = torch.empty(3, dtype=torch.int32, device=counts.device)
    st = torch.empty(3, dtype=torch.float64, device=counts.device)
    check(lib.cgpt_predict_tail(ptr(counts), counts.numel(), float(alpha), ptr(lab), ptr(st),
                                stream_ptr()))
    return lab, st


# ------------------------------------------------------------------------------- model ops
def _dt(t):
    return DT_F32 if t.dtype == torch.float32 else DT_BF16


def norm_rows(x, gamma, beta, eps, out, *, rms=False, rows=None, gather=None):
    """LayerNorm / RMSNorm of rows of x ([M, D], fp32 or bf16) into out ([rows, D])."""
    lib = load()
    D = x.shape[-1]
    rows = out.shape[0] if rows is None else rows
    gp, gs, go = gather if gather is not None else (0, 0, 0)
    check(lib.cgpt_norm_rows(ptr(x), x.stride(0), _dt(x), ptr(gamma), ptr(beta), float(eps), rows, D,
                             ptr(out), out.stride(0), _dt(out), 1 if rms else 0, gp, gs, go, stream_ptr()))
    return out


def attn_vit_supported(*, B, H, T, head_dim):
    """True when a head-major tcgen05 attention kernel serves this non-causal shape: csrc/attn_vit.cu up to 256 (+1 cls)
    keys, csrc/attn_long.cu (multi-tile, exact two-pass softmax) beyond."""
    return 64 < head_dim <= 128 and head_dim % 8 == 0 and T >= 1 and B * H * T < 2 ** 31


def attention(q, k, v, o, *, B, H, Tq, Tk, head_dim, scale, q_rows_per_batch=None,
              kv_rows_per_batch=None, causal=False, kp=None, vp=None, P=0, decode=False, force_flash=False,
              head_major=False):
    """q/k/v/o are 2-D bf16 views whose column 0 is head 0 (e.g. slices of a fused QKV buffer).
    head_major=True: q, k, v are dense [B][H][T][head_dim] blocks (gemm(..., headmajor=...))."""
    lib = load()
    a = AttnArgs()
    a.head_major = 1 if head_major else 0
    if head_major:
        assert q.is_contiguous() and k.is_contiguous() and v.is_contiguous() and Tq == Tk
    a.q = q.data_ptr(); a.ldq = q.stride(0); a.q_rows_per_batch = q_rows_per_batch or Tq
    a.k = k.data_ptr(); a.v = v.data_ptr(); a.ldk = k.stride(0); a.ldv = v.stride(0)
    a.kv_rows_per_batch = kv_rows_per_batch or (Tk - P)
    if P > 0:
        a.kp = kp.data_ptr(); a.vp = vp.data_ptr(); a.ldkp = kp.stride(0); a.ldvp = vp.stride(0)
    a.P = P
    a.o = o.data_ptr(); a.ldo = o.stride(0)
    a.B, a.H, a.Tq, a.Tk, a.head_dim = B, H, Tq, Tk, head_dim
    a.scale = float(scale); a.causal = 1 if causal else 0
    a.decode_kernel = 1 if decode else (2 if force_flash else 0)
    check(lib.cgpt_attention(C.byref(a), stream_ptr()))
    return o


def rope_split(qkv, T, H, head_dim, pos0, cos_t, sin_t, kcache, vcache, cache_rows_per_batch, cache_row0):
    lib = load()
    check(lib.cgpt_rope_split(ptr(qkv), qkv.stride(0), qkv.shape[0], T, H, head_dim, pos0, ptr(cos_t),
                              ptr(sin_t), ptr(kcache), ptr(vcache), kcache.stride(-2),
                              cache_rows_per_batch, cache_row0, stream_ptr()))


def gather_rows(table, ids, rows, out, *, id_period=None, remap=None):
    lib = load()
    D = table.shape[1]
    if id_period is None:
        id_period = ids.numel() if ids is not None else table.shape[0]
    rp, rs, ro = remap if remap is not None else (0, 0, 0)
    check(lib.cgpt_gather_rows(ptr(table), table.stride(0), ptr(ids), id_period, rows, D, ptr(out),
                               out.stride(0), _dt(out), rp, rs, ro, table.shape[0], stream_ptr()))
    return out


def greedy_step(next_idx, finished, ids_out, t, eos_id, pad_id, unfinished_count=None):
    lib = load()
    check(lib.cgpt_greedy_step(ptr(next_idx), next_idx.numel(), ptr(finished), ptr(ids_out),
                               ids_out.stride(0), t, eos_id, pad_id, ptr(unfinished_count), stream_ptr()))


def cosine_rows(feats, target, out=None):
    """scores[r] = cos(feats[r], target), fp32 (CLIP feature scoring of the attack loop)."""
    lib = load()
    assert feats.dtype == torch.float32 and target.dtype == torch.float32 and feats.stride(1) == 1
    rows, D = feats.shape
    assert target.numel() == D and target.is_contiguous()
    if out is None:
        out = torch.empty(rows, dtype=torch.float32, device=feats.device)
    check(lib.cgpt_cosine_rows(ptr(feats), feats.stride(0), rows, D, ptr(target), ptr(out), stream_ptr()))
    return out
