"""CPU: the CLIP vision-tower restatement (oracle/clip_oracle.py) equals transformers' own
CLIPVisionModelWithProjection (the third-party model the reference's attack prose names; the reference ships no
code for this stage), and the attack update rule behaves as specified."""
import pytest
import torch

from certifiedgpt_b200.attack import ClipVisionConfig
from oracle import clip_oracle as co


def _hf(cfg, seed):
    from transformers import CLIPVisionConfig, CLIPVisionModelWithProjection
    torch.manual_seed(seed)
    hc = CLIPVisionConfig(hidden_size=cfg.hidden, intermediate_size=cfg.mlp, num_hidden_layers=cfg.layers,
                          num_attention_heads=cfg.heads, image_size=cfg.img_size, patch_size=cfg.patch,
                          projection_dim=cfg.proj, layer_norm_eps=cfg.eps, hidden_act="quick_gelu")
    m = CLIPVisionModelWithProjection(hc).eval()
    with torch.no_grad():                      # HF leaves biases at 0 and LayerNorms at 1/0: make every term matter
        for n, p in m.named_parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
    return m


@pytest.mark.parametrize("cfg", [ClipVisionConfig.tiny(),
                                 ClipVisionConfig(img_size=84, hidden=96, layers=3, heads=6, mlp=160, proj=40)])
def test_restatement_matches_transformers_clip(cfg):
    m = _hf(cfg, seed=3)
    x = torch.randn(3, 3, cfg.img_size, cfg.img_size, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        ref = m(pixel_values=x).image_embeds
        got = co.clip_vision_features(m.state_dict(), cfg, x)
    assert got.shape == ref.shape
    assert (got - ref).abs().max().item() < 1e-4 * max(1.0, ref.abs().max().item())


def test_rgf_step_moves_along_the_estimated_gradient_and_respects_the_ball():
    g = torch.Generator().manual_seed(0)
    x_clean = torch.rand(3, 8, 8, generator=g)
    u = torch.randn(512, 3, 8, 8, generator=g)
    direction = torch.randn(3, 8, 8, generator=g)
    sigma_q = 0.05
    scores = (u * direction).flatten(1).sum(1) * sigma_q + 0.3        # a linear score: exact finite differences
    x_new, grad = co.rgf_step(x_clean, x_clean, u, 0.3, scores, sigma_q, step_size=0.01, eps=0.02)
    cos = torch.nn.functional.cosine_similarity(grad.flatten(), direction.flatten(), dim=0)
    assert cos > 0.7               # 512 directions in 192 dims: E[cos] ~ sqrt(Q / (Q + d)) = 0.85
    assert (x_new - x_clean).abs().max() <= 0.01 + 1e-7
    assert x_new.min() >= 0 and x_new.max() <= 1
    far = x_clean + 0.02 * torch.sign(direction)
    x2, _ = co.rgf_step(far.clamp(0, 1), x_clean, u, 0.3, scores, sigma_q, step_size=0.01, eps=0.02)
    assert (x2 - x_clean).abs().max() <= 0.02 + 1e-6                   # projection onto the eps-ball
    assert torch.allclose(co.cosine_scores(torch.tensor([[1.0, 0.0], [0.0, 2.0], [3.0, 3.0]]), torch.tensor([1.0, 0.0])),
                          torch.tensor([1.0, 0.0, 0.5 ** 0.5]), atol=1e-6)
