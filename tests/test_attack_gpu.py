"""GPU: the attack loop's kernels and host logic (certifiedgpt_b200/attack.py) against the CPU restatement
(oracle/clip_oracle.py, itself pinned to transformers' CLIP in tests/test_oracle_clip_cpu.py).
Tolerance: bf16 tensor-core path vs fp32 oracle, rel 2e-2 (BASELINE.json north_star)."""
import pytest
import torch

from certifiedgpt_b200.attack import BlackBoxAttack, ClipVisionConfig, ClipVisionEngine
from certifiedgpt_b200.weights import round_to_bf16
from oracle import clip_oracle as co
from oracle import philox_oracle as po

pytestmark = pytest.mark.gpu
REL = 2e-2


def _model(cfg, seed):
    from transformers import CLIPVisionConfig, CLIPVisionModelWithProjection
    torch.manual_seed(seed)
    hc = CLIPVisionConfig(hidden_size=cfg.hidden, intermediate_size=cfg.mlp, num_hidden_layers=cfg.layers,
                          num_attention_heads=cfg.heads, image_size=cfg.img_size, patch_size=cfg.patch,
                          projection_dim=cfg.proj, layer_norm_eps=cfg.eps, hidden_act="quick_gelu")
    m = CLIPVisionModelWithProjection(hc).eval()
    with torch.no_grad():
        for n, p in m.named_parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
    return round_to_bf16({k: v.clone() for k, v in m.state_dict().items()})


def _rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


MID = ClipVisionConfig(img_size=224, hidden=256, layers=3, heads=4, mlp=1024, proj=128)     # hd 64, T = 257 like ViT-L/14


@pytest.mark.parametrize("cfg", [ClipVisionConfig.tiny(), MID])
def test_clip_features_match_oracle(cfg):
    sd = _model(cfg, seed=5)
    eng = ClipVisionEngine(cfg, sd)
    x = torch.randn(4, 3, cfg.img_size, cfg.img_size, generator=torch.Generator().manual_seed(2))
    got = eng.encode_images(x.cuda())
    torch.cuda.synchronize()
    ref = co.clip_vision_features(sd, cfg, x)
    assert _rel(got, ref) < REL
    # cosine scores agree to well below the spread of the scores
    tgt = ref[0]
    from certifiedgpt_b200 import _lib as L
    s_got = L.cosine_rows(got, got[0].clone())
    assert (s_got.cpu() - co.cosine_scores(ref, tgt)).abs().max().item() < 2e-2


def test_cosine_rows_kernel_exact():
    from certifiedgpt_b200 import _lib as L
    g = torch.Generator().manual_seed(0)
    f, t = torch.randn(37, 768, generator=g), torch.randn(768, generator=g)
    got = L.cosine_rows(f.cuda(), t.cuda()).cpu()
    assert (got - co.cosine_scores(f, t)).abs().max().item() < 1e-5


def test_perturbed_queries_use_the_philox_directions_and_pixel_space_normalize():
    """encode_perturbed(x, Q, sigma) == encode_images((x + sigma*u_b - mean)/std) with u_b the Philox draws that
    cgpt_noise_image regenerates for the update (same keys): the estimate is built from the directions the
    encoder actually saw."""
    from certifiedgpt_b200 import _lib as L
    cfg = ClipVisionConfig.tiny()
    sd = _model(cfg, seed=7)
    eng = ClipVisionEngine(cfg, sd)
    S, Q, sigma = cfg.img_size, 6, 0.05
    x = torch.rand(3, S, S, generator=torch.Generator().manual_seed(4)).cuda()
    got = eng.encode_perturbed(x, Q, sigma, seed=9, stream_id=3).clone()
    u = L.noise_image(torch.zeros_like(x), Q, 1.0, seed=9, stream_id=3)
    # the numpy Philox oracle draws the same directions
    import numpy as np
    u_ref = torch.from_numpy(po.draws(3 * S * S, np.arange(Q), seed=9, stream_id=3)).view(Q, 3, S, S)
    assert (u.cpu() - u_ref).abs().max().item() < 1e-4      # device logf / sincospif vs numpy: a few ulp of the draw
    m = torch.tensor(L.BLIP_MEAN, device="cuda").view(1, 3, 1, 1)
    sdv = torch.tensor(L.BLIP_STD, device="cuda").view(1, 3, 1, 1)
    imgs = (x[None] + sigma * u - m) / sdv
    want = eng.encode_images(imgs)
    assert _rel(got, want) < 1e-2          # same kernels; only the bf16 rounding point of the input differs
    ref = co.clip_vision_features(sd, cfg, imgs.cpu())
    assert _rel(got, ref) < REL


def test_attack_loop_matches_oracle_update_and_raises_the_score():
    cfg = ClipVisionConfig.tiny()
    sd = _model(cfg, seed=11)
    eng = ClipVisionEngine(cfg, sd)
    S = cfg.img_size
    g = torch.Generator().manual_seed(8)
    x_clean, target = torch.rand(3, S, S, generator=g), torch.rand(3, S, S, generator=g)
    atk = BlackBoxAttack(eng, None, steps=8, queries=256, sigma_q=0.1, step_size=2.0 / 255, eps=16.0 / 255, seed=1)
    # one step against the oracle's update rule, fed with the engine's own scores and directions
    from certifiedgpt_b200 import _lib as L
    xc = x_clean.cuda()
    tfeat = eng.encode_perturbed(target.cuda(), 1, 0.0)[0].clone()
    x1, f0, scores = atk.step(xc, xc, tfeat, 0)
    u = L.noise_image(torch.zeros_like(xc), 256, 1.0, seed=1, stream_id=0).cpu()
    x1_ref, _ = co.rgf_step(x_clean, x_clean, u, f0.item(), scores.cpu(), 0.1, 2.0 / 255, 16.0 / 255)
    assert (x1.cpu() - x1_ref).abs().max().item() <= 2.0 * 2.0 / 255 * 0.02 + 1e-6 or \
        ((x1.cpu() - x1_ref).abs() > 1e-6).float().mean().item() < 0.02     # sign flips only where |g| ~ 0
    # the whole loop: deterministic, inside the ball, and the CLIP cosine to the target goes up
    adv, info = atk.run(x_clean, target)
    adv2, info2 = atk.run(x_clean, target)
    assert torch.equal(adv, adv2) and info["final_score"] == info2["final_score"]
    assert info["linf"] <= 16.0 / 255 + 1e-6 and len(info["steps"]) == 8
    assert info["final_score"] > info["steps"][0]["score_before"]


def test_attack_with_smoothed_victim_prediction_per_step():
    """configs[4] end to end at toy size: each step also runs Smooth.predict(N) on the perturbed image through the
    native MiniGPT-4 engine (pixel-space noise, Normalize inside K1)."""
    from certifiedgpt_b200.config import ModelConfig
    from certifiedgpt_b200.native import NativeMiniGPT4Engine
    from certifiedgpt_b200.randomized_smoothing.smoothing import Smooth
    from certifiedgpt_b200.weights import random_state_dict
    mcfg = ModelConfig.tiny()
    msd = round_to_bf16(random_state_dict(mcfg, seed=1))
    table = [((t,), t % 5) for t in range(3, mcfg.llm.vocab)]
    victim = NativeMiniGPT4Engine(mcfg, msd, (1, 5, 6), (7, 8, 9, 10), table, 6, max_new_tokens=1)
    smooth = Smooth(victim, 6, 0.25, noise_space="pixel", seed=5)
    ccfg = ClipVisionConfig.tiny()
    eng = ClipVisionEngine(ccfg, _model(ccfg, seed=13))
    S = ccfg.img_size
    assert S == mcfg.vit.img_size
    g = torch.Generator().manual_seed(21)
    atk = BlackBoxAttack(eng, smooth, steps=3, queries=32, predict_n=40, predict_batch=16, seed=2)
    adv, info = atk.run(torch.rand(3, S, S, generator=g), torch.rand(3, S, S, generator=g))
    preds = [s["smoothed_prediction"] for s in info["steps"]]
    assert len(preds) == 3 and all(isinstance(p, int) and -1 <= p < 6 for p in preds)
