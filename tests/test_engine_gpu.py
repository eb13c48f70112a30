"""GPU parity: the MiniGPT-4 engine (sm_100a kernels) vs the fp32 CPU oracle, tower by tower and
end to end through Smooth, on identical weights (bf16-rounded) and identical injected noise.

Tolerances (BASELINE.json north_star): features/logits within bf16 tolerance, rel 2e-2; token
ids / labels / counts bit-exact wherever the oracle's top-2 logit margin exceeds 1e-2."""
import numpy as np
import pytest
import torch

from certifiedgpt_b200.config import LlmConfig, ModelConfig, QFormerConfig, VitConfig
from certifiedgpt_b200.weights import random_state_dict, round_to_bf16
from oracle import model_oracle as mo
from oracle import smoothing_oracle as so

pytestmark = pytest.mark.gpu
REL = 2e-2
MARGIN = 1e-2


def _rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


def _setup(cfg, seed, max_new=3, n_classes=8, prefix=(1, 5, 6), suffix=(7, 8, 9, 10, 11), all_pairs=False):
    from certifiedgpt_b200.engine import MiniGPT4Engine
    sd = round_to_bf16(random_state_dict(cfg, seed=seed))
    V = cfg.llm.vocab
    table = [((t,), t % (n_classes - 1)) for t in range(3, V)]
    if all_pairs:   # every 2-token answer has its own class: label histograms of 2-token runs are not degenerate
        table += [((t, u), (t * 7 + u) % (n_classes - 1)) for t in range(3, V) for u in range(3, V)]
    else:
        table += [((t, u), (t + u) % (n_classes - 1)) for t in range(3, V, 7) for u in range(3, V, 5)]
    eng = MiniGPT4Engine(cfg, sd, prefix, suffix, table, n_classes, max_new_tokens=max_new)
    orc = mo.MiniGPT4ClassifierOracle(sd, cfg, prefix, suffix, table, n_classes, max_new_tokens=max_new)
    return sd, eng, orc


WIDE = ModelConfig(vit=VitConfig(img_size=56, depth=2), qf=QFormerConfig(layers=2),
                   llm=LlmConfig(hidden=512, layers=2, heads=4, inter=1024, vocab=512))


@pytest.mark.parametrize("name,cfg", [("tiny", ModelConfig.tiny()), ("wide", WIDE)])
def test_towers_match_oracle(name, cfg):
    sd, eng, orc = _setup(cfg, seed=21)
    g = torch.Generator().manual_seed(9)
    images = torch.randn(5, 3, cfg.vit.img_size, cfg.vit.img_size, generator=g)
    got = {}
    labels = eng.forward_images(images.cuda(), collect=got)
    torch.cuda.synchronize()
    ref = {}
    with torch.no_grad():
        img = mo.encode_img(sd, cfg, images, collect=ref)
        embeds = mo.build_prompt_embeds(sd, cfg, img, orc.prefix_ids, orc.suffix_ids)
        ids, first_logits, margins = mo.generate_ids(sd, cfg, embeds, orc.max_new_tokens)
    assert _rel(got["embed"], ref["embed"]) < REL
    for i in range(cfg.vit.depth):
        assert _rel(got[f"block{i}"], ref[f"block{i}"]) < REL, f"vit block {i}"
    assert _rel(got["image_embeds"], ref["image_embeds"]) < REL
    for i in range(cfg.qf.layers):
        assert _rel(got[f"layer{i}"], ref[f"layer{i}"]) < REL, f"qformer layer {i}"
    P = len(orc.prefix_ids)
    assert _rel(got["llm_in"], embeds[:, P:]) < REL
    assert _rel(got["first_logits"], first_logits) < REL
    # greedy ids: exact where every step's top-2 margin is safe
    safe = (margins > MARGIN).all(dim=1)
    assert safe.float().mean() > 0.5, "test inputs too close to ties to be informative"
    gids = got["ids"].cpu().long()
    assert torch.equal(gids[safe], ids[safe])
    ref_labels = torch.tensor([mo.answer_label(r.tolist(), orc.table, orc.num_classes - 1) for r in ids])
    assert torch.equal(labels.cpu().long()[safe], ref_labels[safe])


def test_eos_and_padding_semantics():
    """min_length=1 suppresses EOS on the first token; finished rows emit pad (HF greedy search)."""
    cfg = ModelConfig.tiny()
    from certifiedgpt_b200.engine import MiniGPT4Engine
    sd = random_state_dict(cfg, seed=33)
    sd["llama_model.lm_head.weight"][cfg.llm.eos_id] *= 6.0      # EOS strongly preferred
    sd = round_to_bf16(sd)
    table = [((t,), t % 5) for t in range(3, cfg.llm.vocab)]
    eng = MiniGPT4Engine(cfg, sd, (1, 4), (6, 7, 8), table, 6, max_new_tokens=5)
    orc = mo.MiniGPT4ClassifierOracle(sd, cfg, (1, 4), (6, 7, 8), table, 6, max_new_tokens=5)
    images = torch.randn(6, 3, cfg.vit.img_size, cfg.vit.img_size, generator=torch.Generator().manual_seed(2))
    got = {}
    eng.forward_images(images.cuda(), collect=got)
    orc(images)
    ids, margins = orc.last["ids"], orc.last["margins"]
    safe = (margins > MARGIN).all(dim=1)
    gids = got["ids"].cpu().long()
    assert (gids[:, 0] != cfg.llm.eos_id).all()
    assert torch.equal(gids[safe], ids[safe])
    assert (ids == cfg.llm.eos_id).any(), "case must exercise EOS"


@pytest.mark.parametrize("space", ["normalized", "pixel"])
def test_smooth_certify_with_engine_matches_oracle(space):
    from certifiedgpt_b200 import _lib as L
    from certifiedgpt_b200.randomized_smoothing.smoothing import Smooth
    # the wide test model with every 2-token answer a class: ~90 % of the draws are margin-safe on the CPU oracle and
    # they fall into 2-4 classes (the tiny model answers the same whatever the noise: nothing to compare per sample)
    cfg = WIDE
    sd, eng, orc = _setup(cfg, seed=5, max_new=2, n_classes=6, all_pairs=True)
    if space == "pixel":
        orc.normalize = (L.BLIP_MEAN, L.BLIP_STD)
    S = cfg.vit.img_size
    x = torch.rand(3, S, S, generator=torch.Generator().manual_seed(1000))
    if space == "normalized":
        m = torch.tensor(L.BLIP_MEAN).view(3, 1, 1)
        s = torch.tensor(L.BLIP_STD).view(3, 1, 1)
        x = (x - m) / s
    n0, n, sigma, alpha = 40, 200, 0.25, 0.001
    eps = torch.randn(n0 + n, 3, S, S, generator=torch.Generator().manual_seed(1234))
    ours = Smooth(eng, 6, sigma, noise_space=space)
    ours.inject_noise(eps.cuda())
    label, radius = ours.certify(x.cuda(), n0, n, alpha, 64)
    got_sel = ours.last_counts_selection.cpu()
    got_est = ours.last_counts_estimation.cpu()
    assert int(got_sel.sum()) == n0 and int(got_est.sum()) == n
    # per sample: the engine's label of every draw equals the oracle's wherever the oracle's top-2 margin is safe
    kw = dict(noise_space=L.SPACE_PIXEL, mean=L.BLIP_MEAN, std=L.BLIP_STD) if space == "pixel" else {}
    # (.clone(): the Python engine returns a view of its label buffer, which the next call overwrites)
    lab = torch.cat([eng.noisy_labels(x.cuda(), min(50, n0 + n - f), sigma, eps=eps[f:f + 50].cuda(), first_sample=f,
                                      **kw).clone() for f in range(0, n0 + n, 50)]).cpu().long()
    assert torch.equal(torch.bincount(lab[:n0], minlength=6), got_sel)      # counts = histogram of per-sample labels
    assert torch.equal(torch.bincount(lab[n0:], minlength=6), got_est)
    ref, margins = [], []
    for f in range(0, n0 + n, 60):
        orc(x[None] + eps[f:f + 60] * sigma)
        ref.append(orc.last["labels"].clone())
        margins.append(orc.last["margins"].clone())
    ref, safe = torch.cat(ref), (torch.cat(margins) > MARGIN).all(dim=1)
    assert safe.float().mean() > 0.5
    assert torch.equal(lab[safe], ref[safe])
    ref_sel, ref_est = torch.bincount(ref[:n0], minlength=6), torch.bincount(ref[n0:], minlength=6)
    assert int((got_sel - ref_sel).abs().sum()) <= 2 * int((~safe[:n0]).sum())
    assert int((got_est - ref_est).abs().sum()) <= 2 * int((~safe[n0:]).sum())
    if torch.equal(got_sel, ref_sel) and torch.equal(got_est, ref_est):
        assert (label, radius) == so.certify_tail(ref_sel.numpy(), ref_est.numpy(), n, alpha, sigma)


def test_noisy_labels_per_sample_match_oracle_where_margin_safe():
    from certifiedgpt_b200 import _lib as L
    cfg = ModelConfig.tiny()
    sd, eng, orc = _setup(cfg, seed=8, max_new=2, n_classes=6)
    S = cfg.vit.img_size
    x = torch.randn(3, S, S, generator=torch.Generator().manual_seed(4)) * 0.5
    B = 96
    eps = torch.randn(B, 3, S, S, generator=torch.Generator().manual_seed(77))
    labels = eng.noisy_labels(x.cuda(), B, 0.5, eps=eps.cuda()).cpu().long()
    orc(x[None] + eps * 0.5)
    safe = (orc.last["margins"] > MARGIN).all(dim=1)
    assert safe.float().mean() > 0.5
    assert torch.equal(labels[safe], orc.last["labels"][safe])
    # Philox path: same engine, no injected noise, label histogram sums to B
    lab2 = eng.noisy_labels(x.cuda(), B, 0.5, seed=3)
    assert lab2.numel() == B and int(lab2.min()) >= 0 and int(lab2.max()) < 6


@pytest.mark.parametrize("name,cfg", [("tiny", ModelConfig.tiny()),
                                      ("wide", ModelConfig(vit=VitConfig(img_size=56, depth=1), qf=QFormerConfig(layers=2),
                                                           llm=LlmConfig(layers=1, inter=128, vocab=96)))])
def test_image_tower_matches_the_reference_encode_img_run(name, cfg):
    """CUDA image tower (noise-free batch of distinct images) against the run of the reference's OWN MiniGPT4.encode_img
    over its own ViT / Q-Former modules (tests/golden/ref_encode_img.pt, fp32 weights): bf16 tolerance."""
    import os
    from certifiedgpt_b200.engine import MiniGPT4Engine
    ref = torch.load(os.path.join(os.path.dirname(__file__), "golden", "ref_encode_img.pt"))[name]
    sd = random_state_dict(cfg, seed=ref["seed"])
    eng = MiniGPT4Engine(cfg, sd, (1, 5, 6), (7, 8, 9), [((3,), 0)], 2, max_new_tokens=1)
    got = {}
    eng.forward_images(ref["images"].cuda(), collect=got)
    torch.cuda.synchronize()
    tower = got["llm_in"][:, :cfg.qf.n_query].float().cpu()
    assert tower.shape == (3, cfg.qf.n_query, cfg.llm.hidden)
    assert _rel(tower[..., ::ref["stride"]], ref["inputs_llama"]) < REL
