"""Wave quantisation at the per-rank batch of the 4- and 8-GPU runs (275 / 138 draws per rank): every GEMM shape of the
step with 256- vs 128-wide tiles, CTA pairs vs single CTAs.  256x256 pair tiles leave the last wave of the N = 4096
GEMMs (Llama o_proj / down) 43 % full at M = 9 936; this probe measures whether narrower tiles win there.
    python scripts/gemm_wave_probe.py [draws ...]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from certifiedgpt_b200 import _lib as L

draws_list = [int(a) for a in sys.argv[1:]] or [138, 275]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
OPTS = [("auto", 0), ("pair256", 0x2000 | 256), ("pair176", 0x2000 | 176), ("pair128", 0x2000 | 128), ("one256", 0x1000 | 256),
        ("one128", 0x1000 | 128)]
SHAPES = [  # name, rows per draw, N, K, kind
    ("llama_qkv", 72, 12288, 4096, "plain"), ("llama_o", 72, 4096, 4096, "resid"), ("llama_gateup", 72, 22016, 4096, "swiglu"),
    ("llama_down", 72, 4096, 11008, "resid"), ("vit_qkv", 257, 4224, 1408, "plain"), ("vit_proj", 257, 1408, 1408, "resid"),
    ("vit_fc1", 257, 6144, 1408, "gelu"), ("vit_fc2", 257, 1408, 6144, "resid"), ("decode_qkv", 1, 12288, 4096, "plain"),
    ("decode_o", 1, 4096, 4096, "resid"), ("decode_down", 1, 4096, 11008, "resid"), ("decode_gateup", 1, 22016, 4096, "swiglu"), ("lm_head", 1, 32000, 4096, "f32"),
]
if os.environ.get("DECODE_ONLY"):
    SHAPES = [sh for sh in SHAPES if sh[1] == 1]
for draws in draws_list:
    print(f"--- {draws} draws per rank", flush=True)
    for name, rpd, N, K, kind in SHAPES:
        M = draws * rpd
        a = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
        w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
        res = torch.zeros(M, N, device="cuda", dtype=torch.float32) if kind == "resid" else None
        if kind == "swiglu":
            out = torch.empty(M, N // 2, device="cuda", dtype=torch.bfloat16)
        elif kind == "f32":
            out = torch.empty(M, N, device="cuda", dtype=torch.float32)
        else:
            out = None if kind == "resid" else torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        row = []
        for oname, flag in OPTS:
            if M <= 128 and flag & 0x2000:
                row.append(f"{oname}      -")
                continue

            def fn():
                if kind == "resid":
                    L.gemm(a, w, resid=res, out=res, force_bn=flag)
                elif kind == "swiglu":
                    L.gemm(a, w, act=L.ACT_SWIGLU, out=out, force_bn=flag)
                elif kind == "gelu":
                    L.gemm(a, w, act=L.ACT_GELU, out=out, force_bn=flag)
                else:
                    L.gemm(a, w, out=out, force_bn=flag)
            for _ in range(3):
                fn()
            ts = []
            for _ in range(7):
                flush.zero_()
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record(); fn(); e.record()
                torch.cuda.synchronize()
                ts.append(s.elapsed_time(e))
            ms = sorted(ts)[len(ts) // 2]
            row.append(f"{oname} {ms * 1e3:7.1f} us {2.0 * M * N * K / ms / 1e9:5.0f} TF")
        print(f"{name:14s} M={M:6d} N={N:6d} K={K:6d} | " + " | ".join(row), flush=True)
        del a, w, res, out
    torch.cuda.empty_cache()
