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
                                      (20000, 1408, 1408, 0), (130, 1408, 6144, 0), (79, 32000, 4096, 0)])
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


def test_gemm_rejects_bad_shapes(lib):
    a = torch.zeros(8, 12, device="cuda", dtype=torch.bfloat16)
    w = torch.zeros(16, 12, device="cuda", dtype=torch.bfloat16)
    with pytest.raises(lib.CgptError):
        lib.gemm(a, w)
