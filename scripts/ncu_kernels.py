"""Launches every hot-path kernel twice at the full-size shapes of the bench step (batch 1100): the target of
one `ncu --set full` capture (see profiles/r01_ncu_kernels_summary.txt).  First launch = warm, second = the one
summarised."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from certifiedgpt_b200 import _lib as L

B = int(os.environ.get("B", 1100))
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
bf, f32 = torch.bfloat16, torch.float32


def rnd(*shape, dtype=bf, scale=1.0):
    return (torch.randn(*shape, device=dev, generator=g) * scale).to(dtype)


def twice(fn):
    fn()
    if not os.environ.get("ONCE"):      # under ncu one launch is enough (the replay passes warm it)
        fn()
    torch.cuda.synchronize()


T, D, H, hd, MLP = 257, 1408, 16, 88, 6144
Mv = B * T
# K1 noise + normalize + patchify (Philox and injected-noise modes)
x = torch.rand(3, 224, 224, device=dev)
patches = torch.empty(B * 256, 592, device=dev, dtype=bf)
twice(lambda: L.noise_patchify(x, B, 0.25, seed=42, out=patches))
# LayerNorm fp32 -> bf16 (ViT) and RMSNorm (Llama prefill rows)
res = rnd(Mv, D, dtype=f32)
xn = torch.empty(Mv, D, device=dev, dtype=bf)
gam, bet = torch.ones(D, device=dev), torch.zeros(D, device=dev)
twice(lambda: L.norm_rows(res, gam, bet, 1e-6, xn))
# ViT GEMMs: qkv, proj (+residual), fc1 (+GELU), fc2 (+residual)
w_qkv, b_qkv = rnd(3 * D, D, scale=0.02), torch.zeros(3 * D, device=dev)
qkv = torch.empty(Mv, 3 * D, device=dev, dtype=bf)
twice(lambda: L.gemm(xn, w_qkv, bias=b_qkv, out=qkv))
att = torch.empty(Mv, D, device=dev, dtype=bf)
twice(lambda: L.attention(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], att, B=B, H=H, Tq=T, Tk=T, head_dim=hd, scale=hd ** -0.5))
w_proj, b_d = rnd(D, D, scale=0.02), torch.zeros(D, device=dev)
twice(lambda: L.gemm(att, w_proj, bias=b_d, resid=res, out=res))
w_fc1, b_fc1 = rnd(MLP, D, scale=0.02), torch.zeros(MLP, device=dev)
hbuf = torch.empty(Mv, MLP, device=dev, dtype=bf)
twice(lambda: L.gemm(xn, w_fc1, bias=b_fc1, act=L.ACT_GELU, out=hbuf))
w_fc2 = rnd(D, MLP, scale=0.02)
twice(lambda: L.gemm(hbuf, w_fc2, bias=b_d, resid=res, out=res))
del hbuf, qkv, att
# Q-Former: all cross K/V in one GEMM, query-side GEMM, cross attention (flash, hd 64)
w_ckv, b_ckv = rnd(9216, D, scale=0.02), torch.zeros(9216, device=dev)
ckv = torch.empty(Mv, 9216, device=dev, dtype=bf)
twice(lambda: L.gemm(xn, w_ckv, bias=b_ckv, out=ckv))
Mq = B * 32
qh = rnd(Mq, 768)
w_q, b_q = rnd(768, 768, scale=0.02), torch.zeros(768, device=dev)
cq = torch.empty(Mq, 768, device=dev, dtype=bf)
twice(lambda: L.gemm(qh, w_q, bias=b_q, out=cq))
ctx = torch.empty(Mq, 768, device=dev, dtype=bf)
twice(lambda: L.attention(cq, ckv[:, :768], ckv[:, 768:1536], ctx, B=B, H=12, Tq=32, Tk=T, head_dim=64, scale=0.125))
del ckv, res, xn
# Llama prefill: RMSNorm, qkv GEMM, RoPE + cache append, prefill attention, gate/up (SwiGLU), down (+residual)
Hd, I, Tp, P, rows = 4096, 11008, 72, 7, 83
Ml = B * Tp
lres = rnd(Ml, Hd, dtype=f32)
lxn = torch.empty(Ml, Hd, device=dev, dtype=bf)
g1 = torch.ones(Hd, device=dev)
twice(lambda: L.norm_rows(lres, g1, None, 1e-5, lxn, rms=True))
w_lqkv = rnd(3 * Hd, Hd, scale=0.02)
lqkv = torch.empty(Ml, 3 * Hd, device=dev, dtype=bf)
twice(lambda: L.gemm(lxn, w_lqkv, out=lqkv))
kc = torch.zeros(B, rows, Hd, device=dev, dtype=bf)
vc = torch.zeros(B, rows, Hd, device=dev, dtype=bf)
inv = 1.0 / (10000.0 ** (torch.arange(0, 128, 2, dtype=f32) / 128))
fr = torch.outer(torch.arange(128, dtype=f32), inv)
cos_t, sin_t = fr.cos().to(dev), fr.sin().to(dev)
twice(lambda: L.rope_split(lqkv, Tp, 32, 128, P, cos_t, sin_t, kc, vc, rows, P))
latt = torch.empty(Ml, Hd, device=dev, dtype=bf)
twice(lambda: L.attention(lqkv[:, :Hd], kc.view(-1, Hd), vc.view(-1, Hd), latt, B=B, H=32, Tq=Tp, Tk=P + Tp, head_dim=128,
                          scale=128 ** -0.5, kv_rows_per_batch=rows, causal=True))
w_gu = rnd(2 * I, Hd, scale=0.02)
act = torch.empty(Ml, I, device=dev, dtype=bf)
twice(lambda: L.gemm(lxn, w_gu, act=L.ACT_SWIGLU, out=act))
w_down = rnd(Hd, I, scale=0.02)
twice(lambda: L.gemm(act, w_down, resid=lres, out=lres))
del act, lqkv, latt, lres, lxn
# decode step: single-token attention over the KV cache, lm_head GEMM, argmax, labels, histogram, tail
dq = rnd(B, 3 * Hd)
datt = torch.empty(B, Hd, device=dev, dtype=bf)
twice(lambda: L.attention(dq[:, :Hd], kc.view(-1, Hd), vc.view(-1, Hd), datt, B=B, H=32, Tq=1, Tk=P + Tp + 1, head_dim=128,
                          scale=128 ** -0.5, kv_rows_per_batch=rows, causal=True, decode=True))
w_head = rnd(32000, Hd, scale=0.02)
last = rnd(B, Hd)
logits = torch.empty(B, 32000, device=dev, dtype=f32)
twice(lambda: L.gemm(last, w_head, out=logits))
twice(lambda: L.argmax_rows(logits, suppress_col=2, want_margin=True))
labels = torch.randint(0, 3130, (B,), device=dev, dtype=torch.int32)
counts = torch.zeros(3130, dtype=torch.int64, device=dev)
twice(lambda: L.label_hist(labels, counts))
sel = torch.zeros(3130, dtype=torch.int64, device=dev); sel[5] = 100
est = torch.zeros(3130, dtype=torch.int64, device=dev); est[5] = 990; est[6] = 10
twice(lambda: L.certify_tail(sel, est, 1000, 0.001, 0.25))
print("ncu_kernels: done")
