"""GPU parity: tcgen05 GEMM core vs fp32 torch reference of the same op (bf16 inputs)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(a, w, bias=None, act=0, resid=None, row_add=None):
    y = a.float() @ w.float().t()
    if bias is not None:
        y = y + bias
    if act == 1:
        y = torch.nn.functional.gelu(y)  # exact erf GELU (eva_vit.Mlp act_layer=nn.GELU)
    if act == 2:
        y = torch.nn.functional.silu(y[:, 0::2]) * y[:, 1::2]
    if row_add is not None:
        y = y + row_add
    if resid is not None:
        y = y + resid.float()
    return y


def _close(y, r, tol=1.6e-2):
    # bf16 output rounding: rel 2^-8 of the value; fp32 accumulate
    err = (y.float() - r).abs()
    assert (err <= tol * r.abs() + 2e-2 * r.abs().mean()).all(), f"max err {err.max().item()}"


@pytest.mark.parametrize("M,N,K,bn", [(128, 128, 64, 128), (128, 256, 128, 256), (128, 176, 128, 176),
                                      (300, 1408, 592, 0), (771, 4224, 1408, 0), (1, 768, 768, 0),
                                      (20000, 1408, 1408, 0), (130, 1408, 6144, 0), (79, 32000, 4096, 0),
                                      # masked last n-tile of the CTA-pair kernel: 32 / 80 / 208 / 16 valid columns
                                      (700, 1312, 256, 0), (700, 1360, 256, 0), (700, 1488, 256, 0), (700, 1040, 256, 0)])
def test_gemm_plain(lib, M, N, K, bn):
    torch.manual_seed(M + N + K)
    a = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    y = lib.gemm(a, w, force_bn=bn)
    _close(y, _ref(a, w))


def test_gemm_epilogues(lib):
    torch.manual_seed(1)
    M, N, K = 771, 1408, 1408
    a = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    bias = torch.randn(N, device="cuda")
    _close(lib.gemm(a, w, bias=bias), _ref(a, w, bias))
    _close(lib.gemm(a, w, bias=bias, act=lib.ACT_GELU), _ref(a, w, bias, act=1))
    resid = torch.randn(M, N, device="cuda")
    y = lib.gemm(a, w, bias=bias, resid=resid, out_dtype=torch.float32)
    assert torch.allclose(y, _ref(a, w, bias, resid=resid), atol=2e-4, rtol=1e-5)
    # in-place residual update (out aliases resid), as the transformer blocks use it
    r2 = resid.clone()
    lib.gemm(a, w, bias=bias, resid=r2, out=r2)
    assert torch.allclose(r2, y, atol=0, rtol=0)
    rb = resid.bfloat16()
    _close(lib.gemm(a, w, bias=bias, resid=rb), _ref(a, w, bias, resid=rb))
    w2 = (torch.randn(2048, K, device="cuda") * 0.05).bfloat16()
    _close(lib.gemm(a, w2, act=lib.ACT_SWIGLU), _ref(a, w2, act=2), tol=2e-2)


def test_gemm_patch_embed_epilogue(lib):
    """bias + pos_embed[1+p] + scatter of row b*256+p to row b*257+1+p (eva_vit.py:333-340)."""
    torch.manual_seed(2)
    B, P, K, N = 3, 256, 592, 1408
    a = (torch.randn(B * P, K, device="cuda") * 0.5).bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    bias = torch.randn(N, device="cuda")
    pos = torch.randn(P + 1, N, device="cuda")
    out = torch.zeros(B * (P + 1), N, device="cuda")
    lib.gemm(a, w, bias=bias, out=out, row_add=pos, row_period=P, row_add_offset=1,
             remap_stride=P + 1, remap_offset=1)
    ref = (a.float() @ w.float().t() + bias).view(B, P, N) + pos[1:]
    got = out.view(B, P + 1, N)
    assert torch.allclose(got[:, 1:], ref, atol=2e-4, rtol=1e-5)
    assert (got[:, 0] == 0).all()


@pytest.mark.parametrize("M,N,K", [(4096, 512, 4096),      # 8 full 512-row tiles x 2 n-tiles
                                   (5000, 768, 4160),      # ragged last m-tile (392 of 512 rows), ragged K (65 k-blocks of 64)
                                   (4500, 1408, 6144),     # ViT fc2 shape class, masked last n-tile
                                   (4100, 4096, 11008)])   # Llama down: 172 k-blocks; last m-tile holds 4 rows
def test_gemm_512_row_tiles_for_long_k(lib, M, N, K):
    """The 512 x 256 tile per CTA pair (two M sub-tiles per CTA sharing the B tile, both TMEM accumulator stages inside one
    tile; the default for K >= 8192, forced here for every shape): plain, bias + GELU, SwiGLU, bf16 residual and the in-place
    fp32 residual update, and the same bits as the 256-row tiles."""
    import os
    os.environ["CGPT_GEMM_MT"] = "2"
    try:
        _check_512_row_tiles(lib, M, N, K)
    finally:
        os.environ.pop("CGPT_GEMM_MT", None)


def _check_512_row_tiles(lib, M, N, K):
    torch.manual_seed(M + N + K)
    a = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.02).bfloat16()
    bias = torch.randn(N, device="cuda")
    _close(lib.gemm(a, w), _ref(a, w))
    _close(lib.gemm(a, w, bias=bias, act=lib.ACT_GELU), _ref(a, w, bias, act=1))
    _close(lib.gemm(a, w, act=lib.ACT_SWIGLU), _ref(a, w, act=2), tol=2e-2)
    resid = torch.randn(M, N, device="cuda")
    y = lib.gemm(a, w, bias=bias, resid=resid, out_dtype=torch.float32)
    assert torch.allclose(y, _ref(a, w, bias, resid=resid), atol=1e-3, rtol=1e-5)
    r2 = resid.clone()
    lib.gemm(a, w, bias=bias, resid=r2, out=r2)                    # red.global.add epilogue
    assert torch.allclose(r2, y, atol=1e-3, rtol=1e-5)
    rb = resid.bfloat16()
    _close(lib.gemm(a, w, bias=bias, resid=rb), _ref(a, w, bias, resid=rb))
    # same bits whatever the tile and the number of CTAs: every output is accumulated in the same k order
    import os
    y2 = lib.gemm(a, w, bias=bias)
    assert torch.equal(y2, lib.gemm(a, w, bias=bias, max_ctas=2))
    os.environ["CGPT_GEMM_MT"] = "1"                               # the 256-row tiles every other shape uses
    y1 = lib.gemm(a, w, bias=bias)
    os.environ["CGPT_GEMM_MT"] = "2"
    assert torch.equal(y1, y2)


def test_gemm_rejects_bad_shapes(lib):
    a = torch.zeros(8, 12, device="cuda", dtype=torch.bfloat16)
    w = torch.zeros(16, 12, device="cuda", dtype=torch.bfloat16)
    with pytest.raises(lib.CgptError):
        lib.gemm(a, w)


@pytest.mark.parametrize("B,T,heads,pos0,row0,rows", [(3, 72, 4, 7, 7, 83), (40, 1, 4, 80, 80, 83), (5, 7, 32, 0, 0, 7),
                                                      (130, 5, 2, 3, 9, 20),
                                                      (10, 8, 3, 2, 2, 12), (40, 8, 3, 2, 2, 12)])   # odd head count: half-empty last tile
def test_fused_rope_kv_append_epilogue_matches_separate_pass(lib, B, T, heads, pos0, row0, rows):
    """QKV GEMM with the fused rotary + KV-cache-append epilogue vs plain GEMM + cgpt_rope_split and vs an fp32
    torch restatement of HF's apply_rotary_pos_emb (rotate_half).  The fused path rotates the fp32 accumulators
    (one rounding), the separate pass rotates bf16-rounded values (two roundings): compare both to the fp32 truth."""
    hd, K = 128, 256
    D = heads * hd
    g = torch.Generator(device="cuda").manual_seed(T + heads)
    x = torch.randn(B * T, K, device="cuda", generator=g).bfloat16()
    w = (torch.randn(3 * D, K, device="cuda", generator=g) * 0.08).bfloat16()
    inv = 1.0 / (10000.0 ** (torch.arange(0, hd, 2, dtype=torch.float32) / hd))
    fr = torch.outer(torch.arange(256, dtype=torch.float32), inv)
    cos_t, sin_t = fr.cos().cuda().contiguous(), fr.sin().cuda().contiguous()

    def fresh():
        return (torch.full((B, rows, D), 7.0, device="cuda").bfloat16(), torch.full((B, rows, D), -3.0, device="cuda").bfloat16(),
                torch.zeros(B * T, 3 * D, device="cuda", dtype=torch.bfloat16))

    kc1, vc1, q1 = fresh()
    lib.gemm(x, w, out=q1, rope=dict(T=T, heads=heads, pos0=pos0, cos=cos_t, sin=sin_t, kcache=kc1, vcache=vc1,
                                     cache_rows=rows, cache_row0=row0))
    kc2, vc2, q2 = fresh()
    lib.gemm(x, w, out=q2)
    lib.rope_split(q2, T, heads, hd, pos0, cos_t, sin_t, kc2, vc2, rows, row0)
    torch.cuda.synchronize()
    # fp32 truth
    y = (x.float() @ w.float().t()).view(B, T, 3, heads, hd)
    pos = pos0 + torch.arange(T, device="cuda")
    c = torch.cat([cos_t[pos], cos_t[pos]], -1)[None, :, None, :]
    s_ = torch.cat([sin_t[pos], sin_t[pos]], -1)[None, :, None, :]
    rot = lambda t: torch.cat([-t[..., hd // 2:], t[..., :hd // 2]], -1)
    q_ref = y[:, :, 0] * c + rot(y[:, :, 0]) * s_
    k_ref = y[:, :, 1] * c + rot(y[:, :, 1]) * s_
    v_ref = y[:, :, 2]
    scale = max(1.0, y.abs().max().item())
    for got_q, got_k, got_v, tol in [(q1, kc1, vc1, 1e-2), (q2, kc2, vc2, 2e-2)]:
        assert (got_q[:, :D].float().view(B, T, heads, hd) - q_ref).abs().max().item() < tol * scale
        assert (got_k[:, row0:row0 + T].float().view(B, T, heads, hd) - k_ref).abs().max().item() < tol * scale
        assert (got_v[:, row0:row0 + T].float().view(B, T, heads, hd) - v_ref).abs().max().item() < tol * scale
    # cache rows outside [row0, row0 + T) untouched
    keep = torch.ones(rows, dtype=torch.bool)
    keep[row0:row0 + T] = False
    assert (kc1[:, keep] == 7.0).all() and (vc1[:, keep] == -3.0).all()
    # v is a plain copy in both paths: bit-identical
    assert torch.equal(vc1[:, row0:row0 + T], vc2[:, row0:row0 + T])


@pytest.mark.parametrize("M", [1, 100, 138, 275])
@pytest.mark.parametrize("N,K,kind", [(4096, 1024, "resid"), (4096, 4096, "resid"), (22016, 512, "swiglu"), (32000, 256, "f32"),
                                      (1408, 768, "bias"), (4224, 256, "plain")])
def test_skinny_tile_choice_is_bit_identical_to_256_wide_tiles(lib, M, N, K, kind):
    """For one or two m-tiles gemm_bf16 picks the N tile by wave cost (pick_bn_skinny).  The tile width never changes an
    element's K summation order: the automatic choice, forced 256- and forced 128-wide tiles must agree bit for bit
    (this is what keeps labels independent of the batch size), and all match the fp32 reference."""
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = (torch.randn(M, K, device="cuda", generator=g) * 0.5).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).bfloat16()
    pair = 0x2000 if M > 128 else 0x1000
    outs = []
    for flag in (0, pair | 256, pair | 128):
        if kind == "resid":
            r = torch.arange(M * N, device="cuda", dtype=torch.float32).view(M, N) * 1e-6
            lib.gemm(a, w, resid=r, out=r, force_bn=flag)
            outs.append(r)
        elif kind == "swiglu":
            outs.append(lib.gemm(a, w, act=lib.ACT_SWIGLU, force_bn=flag))
        elif kind == "f32":
            outs.append(lib.gemm(a, w, out_dtype=torch.float32, force_bn=flag))
        elif kind == "bias":
            outs.append(lib.gemm(a, w, bias=torch.linspace(-1, 1, N, device="cuda"), force_bn=flag))
        else:
            outs.append(lib.gemm(a, w, force_bn=flag))
    torch.cuda.synchronize()
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    if kind == "resid":
        ref = _ref(a, w) + torch.arange(M * N, device="cuda", dtype=torch.float32).view(M, N) * 1e-6
        assert torch.allclose(outs[0], ref, atol=2e-3, rtol=1e-4)
    elif kind == "swiglu":
        _close(outs[0], _ref(a, w, act=2), tol=2e-2)
    elif kind == "f32":
        assert torch.allclose(outs[0], _ref(a, w), atol=2e-3, rtol=1e-4)
    elif kind == "bias":
        _close(outs[0], _ref(a, w, torch.linspace(-1, 1, N, device="cuda")))
    else:
        _close(outs[0], _ref(a, w))


@pytest.mark.parametrize("B,T,heads,hd", [(3, 257, 16, 88), (5, 17, 4, 88), (2, 130, 3, 96), (1, 64, 2, 128)])
def test_head_major_qkv_scatter_epilogue(lib, B, T, heads, hd):
    """Fused q|k|v projection written as [3][B][heads][T][hd] (what attn_vit.cu reads) = a pure permutation of the
    row-major result: bit-identical values."""
    g = torch.Generator(device="cuda").manual_seed(B * T)
    D = heads * hd
    M = B * T
    a = torch.randn(M, D, device="cuda", generator=g).bfloat16()
    w = (torch.randn(3 * D, D, device="cuda", generator=g) / D ** 0.5).bfloat16()
    bias = torch.randn(3 * D, device="cuda", generator=g)
    plain = lib.gemm(a, w, bias=bias)
    hm = torch.full((M, 3 * D), float("nan"), device="cuda", dtype=torch.bfloat16)
    lib.gemm(a, w, bias=bias, out=hm, headmajor=(T, heads, hd))
    torch.cuda.synchronize()
    want = plain.view(B, T, 3, heads, hd).permute(2, 0, 3, 1, 4).contiguous()
    assert torch.equal(hm.view(3, B, heads, T, hd), want)
    with pytest.raises(lib.CgptError):
        lib.gemm(a, w, bias=bias, out=hm, headmajor=(T + 1, heads, hd))
