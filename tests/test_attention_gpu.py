"""GPU parity: attention kernels (tcgen05 one-shot, mma.sync flash, single-token decode) vs an fp32
torch reference of softmax(scale * Q K^T [+ causal]) V on the same bf16 inputs."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(q, k, v, B, H, Tq, Tk, hd, scale, causal):
    q = q.float().view(B, Tq, H, hd).transpose(1, 2)
    k = k.float().view(B, Tk, H, hd).transpose(1, 2)
    v = v.float().view(B, Tk, H, hd).transpose(1, 2)
    s = (q @ k.transpose(-1, -2)) * scale
    if causal:
        qi = torch.arange(Tq, device=q.device)[:, None] + (Tk - Tq)
        s = s.masked_fill(torch.arange(Tk, device=q.device)[None, :] > qi, float("-inf"))
    return (s.softmax(-1) @ v).transpose(1, 2).reshape(B * Tq, H * hd)


def _run(lib, B, H, Tq, Tk, hd, causal, force_flash, seed=0, fused_qkv=False):
    g = torch.Generator(device="cuda").manual_seed(seed)
    D = H * hd
    if fused_qkv:  # ViT layout: one [B*T, 3D] buffer, attention reads column slices in place
        qkv = (torch.randn(B * Tq, 3 * D, device="cuda", generator=g)).bfloat16()
        q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
    else:
        q = torch.randn(B * Tq, D, device="cuda", generator=g).bfloat16()
        k = torch.randn(B * Tk, D, device="cuda", generator=g).bfloat16()
        v = torch.randn(B * Tk, D, device="cuda", generator=g).bfloat16()
    out = torch.full((B * Tq, D), float("nan"), device="cuda", dtype=torch.bfloat16)
    scale = hd ** -0.5
    lib.attention(q, k, v, out, B=B, H=H, Tq=Tq, Tk=Tk, head_dim=hd, scale=scale, causal=causal,
                  force_flash=force_flash)
    torch.cuda.synchronize()
    ref = _ref(q.contiguous(), k.contiguous(), v.contiguous(), B, H, Tq, Tk, hd, scale, causal)
    err = (out.float() - ref).abs().max().item()
    assert not torch.isnan(out.float()).any(), "unwritten / NaN outputs"
    assert err < 2e-2 * max(1.0, ref.abs().max().item()), f"max err {err}"
    return out


@pytest.mark.parametrize("force_flash", [False, True])
def test_vit_shape_257_tokens_16x88(lib, force_flash):
    _run(lib, 5, 16, 257, 257, 88, False, force_flash, fused_qkv=True)


@pytest.mark.parametrize("force_flash", [False, True])
@pytest.mark.parametrize("Tq,Tk", [(72, 79), (72, 72), (128, 128), (7, 7), (200, 256), (1, 80)])
def test_llama_prefill_shapes_causal_hd128(lib, force_flash, Tq, Tk):
    _run(lib, 4, 8, Tq, Tk, 128, True, force_flash, seed=Tq)


@pytest.mark.parametrize("force_flash", [False, True])
def test_non_causal_hd96_hd104_hd128(lib, force_flash):
    _run(lib, 2, 4, 130, 200, 96, False, force_flash)
    _run(lib, 2, 4, 64, 48, 104, False, force_flash)
    _run(lib, 3, 4, 256, 256, 128, False, force_flash)


@pytest.mark.parametrize("B,H,hd,P,T_own,rows", [(6, 8, 128, 7, 76, 83),     # Llama decode: shared prefix + own rows, ragged tail of 3
                                                   (3, 32, 128, 7, 73, 83),    # full head count, Tk = 80: exactly 10 iterations of 8
                                                   (5, 4, 64, 0, 41, 41),      # no shared prefix, hd = 64
                                                   (2, 4, 32, 3, 2, 9)])       # fewer keys than one iteration
def test_single_token_decode_kernel(lib, B, H, hd, P, T_own, rows):
    """decode_attn_kernel: one query row per sample against [shared prefix rows | the sample's own KV-cache rows]
    (`rows` cache rows are allocated per sample, T_own of them valid)."""
    g = torch.Generator(device="cuda").manual_seed(B * 100 + H + hd)
    D = H * hd
    q = torch.randn(B, D, device="cuda", generator=g).bfloat16()
    kc = torch.randn(B * rows, D, device="cuda", generator=g).bfloat16()
    vc = torch.randn(B * rows, D, device="cuda", generator=g).bfloat16()
    kp = torch.randn(max(P, 1), D, device="cuda", generator=g).bfloat16()
    vp = torch.randn(max(P, 1), D, device="cuda", generator=g).bfloat16()
    out = torch.full((B, D), float("nan"), device="cuda", dtype=torch.bfloat16)
    scale = hd ** -0.5
    lib.attention(q, kc, vc, out, B=B, H=H, Tq=1, Tk=P + T_own, head_dim=hd, scale=scale, q_rows_per_batch=1,
                  kv_rows_per_batch=rows, kp=kp if P else None, vp=vp if P else None, P=P, decode=True)
    torch.cuda.synchronize()
    k_all = torch.cat([kp[:P].float().unsqueeze(0).expand(B, -1, -1), kc.float().view(B, rows, D)[:, :T_own]], 1)
    v_all = torch.cat([vp[:P].float().unsqueeze(0).expand(B, -1, -1), vc.float().view(B, rows, D)[:, :T_own]], 1)
    qh = q.float().view(B, H, 1, hd)
    kh = k_all.view(B, -1, H, hd).transpose(1, 2)
    vh = v_all.view(B, -1, H, hd).transpose(1, 2)
    ref = ((qh @ kh.transpose(-1, -2) * scale).softmax(-1) @ vh).transpose(1, 2).reshape(B, D)
    assert not torch.isnan(out.float()).any()
    err = (out.float() - ref).abs().max().item()
    assert err < 2e-2 * max(1.0, ref.abs().max().item()), f"max err {err}"


def test_flash_only_shapes(lib):
    _run(lib, 3, 12, 32, 32, 64, False, False)          # Q-Former self
    _run(lib, 3, 12, 32, 257, 64, False, False)         # Q-Former cross
    _run(lib, 2, 16, 1025, 1025, 88, False, False)      # 448 px ViT
    _run(lib, 2, 4, 17, 17, 16, False, False)           # tiny test model


def test_umma_and_flash_agree(lib):
    a = _run(lib, 3, 16, 257, 257, 88, False, False, seed=5, fused_qkv=True)
    b = _run(lib, 3, 16, 257, 257, 88, False, True, seed=5, fused_qkv=True)
    assert (a.float() - b.float()).abs().max().item() < 2e-2


@pytest.mark.parametrize("B,H,Tq,Tk,rows,causal", [
    (40, 32, 72, 79, 83, True),      # the smoothing path's prefill: 1280 items -> ~9 per persistent CTA
    (300, 4, 72, 79, 83, True),
    (37, 8, 128, 128, 128, True),    # largest shape: a single load stage (no S-ahead issue)
    (50, 8, 5, 16, 20, True),        # tiny tiles: the 128-row UMMA operand reads past the sub-tiles
    (33, 8, 64, 48, 48, False),
    (1, 32, 7, 7, 7, True),          # the shared-prefix pass
])
def test_persistent_prefill_kernel_kv_cache_layout(lib, B, H, Tq, Tk, rows, causal):
    """attn_prefill.cu: many (sample, head) items per CTA through the load ring / two TMEM stages, keys read
    out of a KV cache whose rows beyond Tk hold finite garbage (must be masked, not multiplied in)."""
    hd, D = 128, H * 128
    g = torch.Generator(device="cuda").manual_seed(B + Tq)
    q = torch.randn(B * Tq, D, device="cuda", generator=g).bfloat16()
    kc = torch.full((B, rows, D), 50.0, device="cuda").bfloat16()
    vc = torch.full((B, rows, D), -70.0, device="cuda").bfloat16()
    kc[:, :Tk] = torch.randn(B, Tk, D, device="cuda", generator=g).bfloat16()
    vc[:, :Tk] = torch.randn(B, Tk, D, device="cuda", generator=g).bfloat16()
    out = torch.full((B * Tq, D), float("nan"), device="cuda", dtype=torch.bfloat16)
    scale = hd ** -0.5
    lib.attention(q, kc.view(-1, D), vc.view(-1, D), out, B=B, H=H, Tq=Tq, Tk=Tk, head_dim=hd, scale=scale,
                  causal=causal, kv_rows_per_batch=rows)
    torch.cuda.synchronize()
    ref = _ref(q, kc[:, :Tk].reshape(-1, D), vc[:, :Tk].reshape(-1, D), B, H, Tq, Tk, hd, scale, causal)
    assert not torch.isnan(out.float()).any(), "unwritten / NaN outputs"
    err = (out.float() - ref).abs().max().item()
    assert err < 2e-2 * max(1.0, ref.abs().max().item()), f"max err {err}"
    # same inputs through the mma.sync flash kernel: independent implementation, same answer
    out2 = torch.empty_like(out)
    lib.attention(q, kc.view(-1, D), vc.view(-1, D), out2, B=B, H=H, Tq=Tq, Tk=Tk, head_dim=hd, scale=scale,
                  causal=causal, kv_rows_per_batch=rows, force_flash=True)
    torch.cuda.synchronize()
    assert (out.float() - out2.float()).abs().max().item() < 2e-2 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("B,H,Tq,Tk,hd,causal,fused", [
    (40, 16, 257, 257, 88, False, True),     # ViT: 640 items -> 4-5 per persistent CTA, cls row on CUDA cores
    (100, 4, 130, 200, 96, False, False),
    (70, 4, 64, 48, 104, False, False),      # one q-tile, Q pad columns stay zero across items
    (80, 8, 200, 256, 128, True, False),
])
def test_persistent_one_shot_kernel_many_items(lib, B, H, Tq, Tk, hd, causal, fused):
    """attn_umma.cu walks several (sample, head) items per CTA: barrier parities, x-buffer double buffering,
    Q pad re-zeroing and the load of item i+1 issued under the epilogue of item i."""
    a = _run(lib, B, H, Tq, Tk, hd, causal, False, seed=3, fused_qkv=fused)
    b = _run(lib, B, H, Tq, Tk, hd, causal, True, seed=3, fused_qkv=fused)
    assert (a.float() - b.float()).abs().max().item() < 2e-2


# ------------------------------------------------------------------ head-major layout + pipelined tcgen05 kernel (attn_vit.cu)
def _run_head_major(lib, B, H, T, hd, seed=0):
    """q, k, v as dense [B][H][T][hd] blocks; reference = the same fp32 softmax attention."""
    g = torch.Generator(device="cuda").manual_seed(seed)
    q, k, v = (torch.randn(B, H, T, hd, device="cuda", generator=g).bfloat16() for _ in range(3))
    out = torch.full((B * T, H * hd), float("nan"), device="cuda", dtype=torch.bfloat16)
    scale = hd ** -0.5
    assert lib.attn_vit_supported(B=B, H=H, T=T, head_dim=hd)
    lib.attention(q.view(-1), k.view(-1), v.view(-1), out, B=B, H=H, Tq=T, Tk=T, head_dim=hd, scale=scale, head_major=True)
    torch.cuda.synchronize()
    s = (q.float() @ k.float().transpose(-1, -2)) * scale
    ref = (s.softmax(-1) @ v.float()).transpose(1, 2).reshape(B * T, H * hd)
    assert not torch.isnan(out.float()).any(), "unwritten / NaN outputs"
    err = (out.float() - ref).abs().max().item()
    assert err < 2e-2 * max(1.0, ref.abs().max().item()), f"max err {err}"
    return out, (q, k, v)


@pytest.mark.parametrize("B,H,T,hd", [
    (5, 16, 257, 88),       # EVA ViT-g at 224 px: cls key / query on CUDA cores, two query tiles
    (40, 16, 257, 88),      # 640 items: every persistent CTA pipelines several items
    (3, 4, 17, 88),         # the small test models (img 56): one tile, 16 -> 32 padded keys masked
    (9, 4, 100, 88),        # one query tile, 112 padded keys
    (7, 3, 200, 96),        # two query tiles, no cls row, hd = 96
    (6, 2, 256, 128),       # full tiles, hd = 128 (O fills the region)
    (4, 5, 129, 72),        # second query tile holds a single row
    (1, 1, 257, 88),        # a single item
])
def test_head_major_pipelined_kernel(lib, B, H, T, hd):
    _run_head_major(lib, B, H, T, hd, seed=T + hd)


def test_head_major_kernel_is_deterministic_and_matches_the_row_major_kernel(lib):
    """Same inputs through the round-1 one-shot kernel (row-major fused QKV) and the pipelined head-major kernel."""
    B, H, T, hd = 12, 16, 257, 88
    a, (q, k, v) = _run_head_major(lib, B, H, T, hd, seed=3)
    b, _ = _run_head_major(lib, B, H, T, hd, seed=3)
    assert torch.equal(a, b)
    D = H * hd
    qkv = torch.cat([t.transpose(1, 2).reshape(B * T, D) for t in (q, k, v)], dim=1).contiguous()
    out = torch.empty(B * T, D, device="cuda", dtype=torch.bfloat16)
    lib.attention(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], out, B=B, H=H, Tq=T, Tk=T, head_dim=hd, scale=hd ** -0.5)
    assert (a.float() - out.float()).abs().max().item() < 1e-2


def test_head_major_rejects_unsupported_shapes(lib):
    q = torch.zeros(2 * 4 * 300 * 88, device="cuda", dtype=torch.bfloat16)
    out = torch.empty(2 * 300, 4 * 88, device="cuda", dtype=torch.bfloat16)
    with pytest.raises(lib.CgptError):
        lib.attention(q[:2 * 4 * 300 * 64], q, q, out, B=2, H=4, Tq=300, Tk=300, head_dim=64, scale=1.0, head_major=True)
    with pytest.raises(lib.CgptError):
        lib.attention(q, q, q, out, B=2, H=4, Tq=100, Tk=100, head_dim=88, scale=1.0, head_major=True, causal=True)


# ------------------------------------------------------------------ multi-tile kernel (attn_long.cu): 448 px ViT, T = 1025
@pytest.mark.parametrize("B,H,T,hd", [
    (2, 16, 1025, 88),      # EVA ViT-g at 448 px: 9 query tiles x 9 key tiles, the last ones hold one row
    (1, 2, 1025, 88),
    (3, 4, 300, 88),        # 3 tiles, ragged last tile (44 rows)
    (2, 3, 258, 96),        # just past the one-tile kernel's range
    (2, 2, 640, 128),       # 5 full tiles, hd = 128
    (5, 4, 513, 72),        # odd tile count, one-row last tile, hd = 72
    (2, 3, 360, 88),        # last tile of 104 keys: both 64-key halves exist, the second one ragged (40 keys)
    (1, 2, 512, 88),        # even tile count: both softmax groups busy in every unit
    (2, 2, 321, 80),        # last tile of 65 keys: the second half holds one key
])
def test_head_major_multi_tile_kernel(lib, B, H, T, hd):
    _run_head_major(lib, B, H, T, hd, seed=T + hd)


@pytest.mark.parametrize("where", ["first", "middle", "last"])
def test_multi_tile_kernel_rescales_when_the_row_maximum_jumps(lib, where):
    """The multi-tile kernel runs an online softmax with a lazy rescale of O (attn_long.cu): plant keys whose scores tower
    over everything before them so that the running maximum jumps by far more than the rescale threshold at a chosen key
    tile - at the very end the previously accumulated O is scaled by ~2^-60 and the result is the planted key's value."""
    B, H, T, hd = 2, 3, 1025, 88
    g = torch.Generator(device="cuda").manual_seed(11)
    q, k, v = (torch.randn(B, H, T, hd, device="cuda", generator=g).bfloat16() for _ in range(3))
    pos = {"first": 5, "middle": 600, "last": T - 1}[where]
    # key `pos` is aligned with every query of head 1 (score ~ +24, the rest stay within +-3), and mildly with those of head 2
    k[:, 1, pos] = 4.0
    q[:, 1] = (q[:, 1].float().abs() * 0.5 + 0.25).bfloat16()
    k[:, 2, pos] = (k[:, 2, pos].float() * 4).bfloat16()
    out = torch.full((B * T, H * hd), float("nan"), device="cuda", dtype=torch.bfloat16)
    scale = hd ** -0.5
    lib.attention(q.view(-1), k.view(-1), v.view(-1), out, B=B, H=H, Tq=T, Tk=T, head_dim=hd, scale=scale, head_major=True)
    torch.cuda.synchronize()
    s = (q.float() @ k.float().transpose(-1, -2)) * scale
    assert (s[:, 1, :, pos] - s[:, 1].amax(-1)).abs().max().item() == 0.0     # the planted key holds every row's maximum
    assert (s[:, 1, :, pos] - s[:, 1, :, :pos].amax(-1)).min().item() > 8.0        # a jump of > 11 in log2 units
    ref = (s.softmax(-1) @ v.float()).transpose(1, 2).reshape(B * T, H * hd)
    assert not torch.isnan(out.float()).any()
    err = (out.float() - ref).abs().max().item()
    assert err < 2e-2 * max(1.0, ref.abs().max().item()), f"max err {err}"


def test_multi_tile_kernel_many_units_and_determinism(lib):
    a, _ = _run_head_major(lib, 12, 16, 1025, 88, seed=9)      # 1728 units: ~12 per persistent CTA
    b, _ = _run_head_major(lib, 12, 16, 1025, 88, seed=9)
    assert torch.equal(a, b)
