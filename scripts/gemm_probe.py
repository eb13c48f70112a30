"""GPU probe for the tcgen05 GEMM: correctness across shapes/epilogues, then throughput.
Run under gpurun; prints one line per case and never stops at the first failure."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from certifiedgpt_b200 import _lib as L

torch.manual_seed(0)
dev = "cuda"

def ref(a, w, bias=None, act=0, resid=None, row_add=None):
    y = a.float() @ w.float().t()
    if bias is not None: y = y + bias
    if act == 1: y = torch.nn.functional.gelu(y)
    if act == 2: y = torch.nn.functional.silu(y[:, 0::2]) * y[:, 1::2]
    if row_add is not None: y = y + row_add
    if resid is not None: y = y + resid.float()
    return y

def case(M, N, K, bn=0, **kw):
    a = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    bias = torch.randn(N, device=dev) if kw.get("bias") else None
    act = kw.get("act", 0)
    resid = None
    if kw.get("resid") == "f32": resid = torch.randn(M, N, device=dev)
    if kw.get("resid") == "bf16": resid = torch.randn(M, N, device=dev).bfloat16()
    odt = torch.float32 if kw.get("f32out") else torch.bfloat16
    try:
        y = L.gemm(a, w, bias=bias, act=act, resid=resid, out_dtype=odt, force_bn=bn)
        torch.cuda.synchronize()
        r = ref(a, w, bias, act, resid)
        err = (y.float() - r).abs().max().item()
        scale = r.abs().max().item()
        ok = err <= 2e-2 * scale + 1e-3
        print(f"{'OK ' if ok else 'BAD'} M={M} N={N} K={K} bn={bn} {kw} maxerr={err:.4g} scale={scale:.3g}", flush=True)
        if not ok:
            d = (y.float() - r).abs()
            bad = (d > 2e-2 * scale + 1e-3)
            rows = bad.any(1).nonzero().flatten()[:8].tolist()
            cols = bad.any(0).nonzero().flatten()[:8].tolist()
            print("   bad rows", rows, "bad cols", cols, "frac", bad.float().mean().item(), flush=True)
        return ok
    except Exception as e:
        print(f"EXC M={M} N={N} K={K} bn={bn} {kw}: {e}", flush=True)
        return False

def perf(M, N, K, bn=0, iters=20, **kw):
    a = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    bias = torch.randn(N, device=dev)
    act = kw.get("act", 0)
    out = torch.empty(M, N // 2 if act == 2 else N, device=dev, dtype=torch.bfloat16)
    for _ in range(3): L.gemm(a, w, bias=bias, act=act, out=out, force_bn=bn)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); s.record()
    for _ in range(iters): L.gemm(a, w, bias=bias, act=act, out=out, force_bn=bn)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / iters
    tf = 2.0 * M * N * K / ms / 1e9
    # cuBLAS reference
    for _ in range(3): torch.matmul(a, w.t())
    torch.cuda.synchronize(); s.record()
    for _ in range(iters): torch.matmul(a, w.t())
    e.record(); torch.cuda.synchronize()
    ms2 = s.elapsed_time(e) / iters
    print(f"PERF M={M} N={N} K={K} bn={bn} act={act}: {ms:.3f} ms {tf:.1f} TF/s | cublas {ms2:.3f} ms {2.0*M*N*K/ms2/1e9:.1f} TF/s", flush=True)

if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), flush=True)
    oks = []
    C1, C2 = 0x1000, 0x2000
    print("--- CTA-pair (cta_group::2) kernels", flush=True)
    oks.append(case(256, 128, 64, 128 | C2))
    oks.append(case(256, 256, 128, 256 | C2))
    oks.append(case(256, 176, 256, 176 | C2))
    oks.append(case(300, 1408, 592, C2))
    oks.append(case(257 * 3, 4224, 1408, C2, bias=True))
    oks.append(case(100, 256, 128, C2))
    oks.append(case(20000, 6144, 1408, C2, bias=True, act=1))
    oks.append(case(5000, 1408, 6144, C2, bias=True, resid="f32", f32out=True))
    print("--- 1-CTA kernels", flush=True)
    oks.append(case(128, 128, 64, 128 | C1))
    oks.append(case(128, 128, 256, 128))
    oks.append(case(128, 256, 128, 256))
    oks.append(case(128, 176, 128, 176))
    oks.append(case(256, 512, 512))
    oks.append(case(300, 1408, 592))          # M tail + K tail (patch embed)
    oks.append(case(257 * 3, 4224, 1408, bias=True))
    oks.append(case(257 * 3, 6144, 1408, bias=True, act=1))
    oks.append(case(257 * 3, 1408, 6144, bias=True, resid="f32", f32out=True))
    oks.append(case(500, 768, 768, bias=True, resid="bf16"))
    oks.append(case(640, 2048, 512, act=2))
    oks.append(case(96, 32000, 4096, f32out=True))
    oks.append(case(20000, 1408, 1408, bias=True))
    print("ALL_OK" if all(oks) else "SOME_BAD", flush=True)
    if "--perf" in sys.argv and oks[0]:
        B = 256
        for flag in (0x1000, 0x2000):
            print("--- perf with", "1-CTA" if flag == 0x1000 else "CTA pairs", flush=True)
            perf(B * 257, 4224, 1408, bn=flag)
            perf(B * 257, 4224, 1408, bn=256 | flag)
            perf(B * 257, 1408, 1408, bn=flag)
            perf(B * 257, 1408, 1408, bn=256 | flag)
            perf(B * 257, 6144, 1408, bn=flag, act=1)
            perf(B * 257, 1408, 6144, bn=flag)
            perf(8192, 8192, 8192, bn=flag)
            perf(B * 72, 12288, 4096, bn=flag)
            perf(B * 72, 22016, 4096, bn=flag, act=2)
