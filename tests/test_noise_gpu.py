"""GPU parity: K1 fused noise/normalise/patchify kernel vs the numpy oracle."""
import numpy as np
import pytest
import torch

from oracle import philox_oracle as po

pytestmark = pytest.mark.gpu


def _bf16(a):
    return torch.from_numpy(np.ascontiguousarray(a)).bfloat16().float().numpy()


@pytest.mark.parametrize("space", ["normalized", "pixel"])
@pytest.mark.parametrize("S", [224, 448])
def test_injected_noise_bit_exact(lib, space, S):
    g = torch.Generator().manual_seed(1000)
    x = torch.rand(3, S, S, generator=g)
    B = 3
    eps = torch.randn(B, 3, S, S, generator=torch.Generator().manual_seed(1234))
    sp = lib.SPACE_PIXEL if space == "pixel" else lib.SPACE_NORMALIZED
    out = lib.noise_patchify(x.cuda(), B, 0.25, eps=eps.cuda(), noise_space=sp)
    torch.cuda.synchronize()
    ms = (lib.BLIP_MEAN, lib.BLIP_STD) if space == "pixel" else (None, None)
    ref = po.patchify(po.noisy_batch(x.numpy(), eps.numpy(), 0.25, *ms))
    got = out.float().cpu().numpy()
    assert got.shape == ref.shape
    assert np.array_equal(got, _bf16(ref))  # bit-exact after the bf16 rounding both sides apply
    assert (got[:, 588:] == 0).all()


def test_philox_mode_matches_oracle_stream(lib):
    S = 224
    x = torch.rand(3, S, S, generator=torch.Generator().manual_seed(1))
    B = 4
    img = lib.noise_image(x.cuda(), B, 0.5, seed=42, stream_id=7, first_sample=10)
    eps = po.draws(3 * S * S, np.arange(10, 10 + B), seed=42, stream_id=7).reshape(B, 3, S, S)
    ref = po.noisy_batch(x.numpy(), eps, 0.5)
    got = img.cpu().numpy()
    # device logf/sincospif vs numpy differ by a few ulp of the draw
    assert np.max(np.abs(got - ref)) < 2e-5
    # patchified bf16 output uses the same draws
    pat = lib.noise_patchify(x.cuda(), B, 0.5, seed=42, stream_id=7, first_sample=10).float().cpu().numpy()
    refp = _bf16(po.patchify(ref))
    assert np.mean(pat != refp) < 2e-3  # only bf16 rounding boundaries may flip
    assert np.max(np.abs(pat - refp)) < 0.05


def test_sample_draw_independent_of_batching(lib):
    x = torch.rand(3, 224, 224, generator=torch.Generator().manual_seed(3)).cuda()
    full = lib.noise_patchify(x, 6, 0.25, seed=5, first_sample=0)
    part = lib.noise_patchify(x, 2, 0.25, seed=5, first_sample=4)
    assert torch.equal(full[4 * 256:], part)
    other = lib.noise_patchify(x, 2, 0.25, seed=5, first_sample=4, stream_id=1)
    assert not torch.equal(other, part)


def test_uniform_mode(lib):
    x = torch.zeros(3, 224, 224).cuda()
    img = lib.noise_image(x, 2, 1.0, seed=9, noise_kind=lib.NOISE_UNIFORM)
    ref = po.draws(3 * 224 * 224, [0, 1], seed=9, kind="uniform").reshape(2, 3, 224, 224)
    assert np.array_equal(img.cpu().numpy(), ref)
    assert 0.49 < img.mean().item() < 0.51


def test_gaussian_moments_full_size(lib):
    x = torch.zeros(3, 224, 224).cuda()
    img = lib.noise_image(x, 64, 1.0, seed=11)
    assert abs(img.mean().item()) < 1e-3 and abs(img.std().item() - 1) < 1e-3
    assert 4.5 < img.abs().max().item() < 7.0


def test_generic_image_shapes(lib):
    x = torch.rand(3, 32, 32).cuda()
    eps = torch.randn(5, 3, 32, 32).cuda()
    out = lib.noise_image(x, 5, 0.3, eps=eps)
    assert torch.equal(out, x[None] + eps * 0.3)
    with pytest.raises(lib.CgptError):
        lib.noise_image(torch.rand(3, 3, 3).cuda(), 1, 0.1)
