"""GPU: the certify / predict agents (empty files in the reference) drive Smooth over a dataset
and log Cohen-style records; results agree with calling Smooth directly."""
import pytest
import torch

from certifiedgpt_b200.config import ModelConfig
from certifiedgpt_b200.weights import random_state_dict, round_to_bf16

pytestmark = pytest.mark.gpu


class _Items:
    def __init__(self, n, size):
        self.items = [{"image": torch.rand(3, size, size, generator=torch.Generator().manual_seed(1000 + i)),
                       "label": i % 5} for i in range(n)]

    def __len__(self):
        return len(self.items)

    def __getitem__(self, i):
        return self.items[i]


def _engine():
    from certifiedgpt_b200.engine import MiniGPT4Engine
    cfg = ModelConfig.tiny()
    sd = round_to_bf16(random_state_dict(cfg, seed=41))
    table = [((t,), t % 5) for t in range(3, cfg.llm.vocab)]
    return cfg, MiniGPT4Engine(cfg, sd, (1, 4, 5), (6, 7, 8, 9), table, 6, max_new_tokens=2)


def test_certify_agent_logs_and_matches_direct_calls(tmp_path):
    from certifiedgpt_b200.agents import setup_agent
    from certifiedgpt_b200.randomized_smoothing.smoothing import Smooth
    cfg, eng = _engine()
    data = _Items(4, cfg.vit.img_size)
    out = tmp_path / "certify.log"
    agent = setup_agent("image_text_certify", base_classifier=eng, dataset=data, num_classes=6, sigma=0.25,
                        n0=20, n=100, alpha=0.001, batch_size=64, outfile=str(out), smooth_kwargs={"seed": 3})
    recs = agent.run()
    agent.finalize()
    assert len(recs) == 4 and all(set(r) >= {"idx", "label", "predict", "radius", "correct", "time"} for r in recs)
    lines = out.read_text().strip().splitlines()
    assert lines[0].split("\t") == ["idx", "label", "predict", "radius", "correct", "time"] and len(lines) == 5
    # same seed + image id -> same answer when Smooth is called directly (order independent)
    direct = Smooth(eng, 6, 0.25, seed=3)
    direct.image_id = 2
    assert direct.certify(data[2]["image"].cuda(), 20, 100, 0.001, 64) == (recs[2]["predict"], recs[2]["radius"])
    acc = agent.certified_accuracy([0.0, 0.5])
    assert 0.0 <= acc[0.5] <= acc[0.0] <= 1.0


def test_predict_agent(tmp_path):
    from certifiedgpt_b200.agents import setup_agent
    from certifiedgpt_b200.randomized_smoothing.smoothing import Smooth
    cfg, eng = _engine()
    data = _Items(3, cfg.vit.img_size)
    agent = setup_agent("image_text_predict", base_classifier=eng, dataset=data, num_classes=6, sigma=0.5,
                        n=64, alpha=0.001, batch_size=32, outfile=str(tmp_path / "p.log"), max_items=2)
    recs = agent.run()
    assert len(recs) == 2 and all(r["predict"] in (Smooth.ABSTAIN, 0, 1, 2, 3, 4, 5) for r in recs)
    assert 0.0 <= agent.abstain_rate() <= 1.0


def test_agents_certify_every_item_under_its_own_question():
    """Items carry their own question (data/vqav2.py::certify_items with a tokenizer): the agent hands it to the engine
    before certifying the item, so the record equals a direct call on an engine built with that question."""
    from certifiedgpt_b200.agents import setup_agent
    from certifiedgpt_b200.engine import MiniGPT4Engine
    from certifiedgpt_b200.native import NativeMiniGPT4Engine
    from certifiedgpt_b200.randomized_smoothing.smoothing import Smooth
    cfg = ModelConfig.tiny()
    sd = round_to_bf16(random_state_dict(cfg, seed=41))
    table = [((t,), t % 5) for t in range(3, cfg.llm.vocab)]
    questions = [(6, 7, 8, 9, 10), (11, 12, 13), (6, 7, 8, 9, 10), (20, 21, 22, 23)]
    eng = NativeMiniGPT4Engine(cfg, sd, (1, 4, 5), max(questions, key=len), table, 6, max_new_tokens=2)
    data = _Items(4, cfg.vit.img_size)
    for it, q in zip(data.items, questions):
        it["suffix_ids"] = list(q)
    agent = setup_agent("image_text_certify", base_classifier=eng, dataset=data, num_classes=6, sigma=0.25,
                        n0=16, n=64, alpha=0.001, batch_size=40, smooth_kwargs={"seed": 3})
    recs = agent.run()
    for i, q in enumerate(questions):
        ref = Smooth(MiniGPT4Engine(cfg, sd, (1, 4, 5), q, table, 6, max_new_tokens=2), 6, 0.25, seed=3)
        ref.image_id = i
        assert ref.certify(data[i]["image"].cuda(), 16, 64, 0.001, 40) == (recs[i]["predict"], recs[i]["radius"]), i
