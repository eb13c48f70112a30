"""Shared by tests/golden/make_ref_generate_fixtures.py (runs the REFERENCE's MiniGPTBase.generate) and
tests/test_oracle_model_cpu.py (runs the oracle on the same inputs): cases, the deterministic character tokenizer and the
transformers Llama built from a certifiedgpt_b200 state dict."""
import torch

PROMPT = "[INST] <Img><ImageHere></Img> [vqa] {} [/INST]"        # conv template + vqav2_dataset.py:39-42,149

CASES = [
    dict(name="single", seed=3, B=1, texts=[PROMPT.format("what color is the car ?")], max_new_tokens=6, eos_boost=0),
    dict(name="batch_same_text", seed=4, B=3, texts=[PROMPT.format("is it raining ?")] * 3, max_new_tokens=5, eos_boost=0),
    dict(name="early_eos", seed=5, B=4, texts=[PROMPT.format("how many dogs ?")] * 4, max_new_tokens=8, eos_boost=3.0),
    dict(name="one_token", seed=6, B=2, texts=[PROMPT.format("what is this ?")] * 2, max_new_tokens=1, eos_boost=0),
    # different questions in one batch: the reference left-pads the shorter prompt (minigpt_base.py:399-412)
    dict(name="ragged_texts", seed=7, B=2, texts=[PROMPT.format("is it raining ?"), PROMPT.format("what color is the big car ?")],
         max_new_tokens=4, eos_boost=0),
]


def encode(text, vocab):
    """character tokenizer: ids in [3, vocab)"""
    return [3 + (ord(c) % (vocab - 3)) for c in text]


SPECIAL = {"<unk>": 0, "<s>": 1, "</s>": 2}


def encode_special(text, vocab):
    """encode() with the literal strings <s>, </s>, <unk> mapped to their ids, as LlamaTokenizer does"""
    import re
    ids = []
    for piece in re.split(r"(<s>|</s>|<unk>)", text):
        if piece in SPECIAL:
            ids.append(SPECIAL[piece])
        elif piece:
            ids.extend(encode(piece, vocab))
    return ids


class _Enc:
    def __init__(self, rows, pad_id):
        n = max(len(r) for r in rows)
        self.input_ids = torch.tensor([r + [pad_id] * (n - len(r)) for r in rows], dtype=torch.long)
        self.attention_mask = torch.tensor([[1] * len(r) + [0] * (n - len(r)) for r in rows], dtype=torch.long)

    def to(self, device):
        return self


class CharTokenizer:
    """The part of LlamaTokenizer the reference's generate and training-forward paths touch (minigpt_base.py:79-82,
    120-131,299-306,441): single strings or right-padded batches, BOS only with add_special_tokens, the literal
    "</s>" of end_sym mapped to the EOS id."""
    padding_side = "right"
    pad_token_id, bos_token_id, eos_token_id = 0, 1, 2
    bos_token = "<s>"

    def __init__(self, vocab):
        self.vocab = vocab

    def __call__(self, text, return_tensors="pt", add_special_tokens=True, padding=False, truncation=False,
                 max_length=None, **kw):
        rows = []
        for t in ([text] if isinstance(text, str) else list(text)):
            ids = ([1] if add_special_tokens else []) + encode_special(t, self.vocab)
            rows.append(ids[:max_length] if truncation and max_length else ids)
        return _Enc(rows, self.pad_token_id)

    def decode(self, ids, skip_special_tokens=False):
        names = {0: "<unk>", 1: "<s>", 2: "</s>"}
        words = []
        for t in [int(v) for v in ids]:
            if t in names:
                if not skip_special_tokens:
                    words.append(names[t])
            else:
                words.append(f"t{t}")
        return " ".join(words)


def hf_llama(cfg, sd):
    from transformers import LlamaConfig, LlamaForCausalLM
    l = cfg.llm
    hc = LlamaConfig(hidden_size=l.hidden, intermediate_size=l.inter, num_hidden_layers=l.layers,
                     num_attention_heads=l.heads, num_key_value_heads=l.heads, vocab_size=l.vocab,
                     rms_norm_eps=l.rms_eps, rope_theta=l.rope_theta, max_position_embeddings=256,
                     bos_token_id=1, eos_token_id=l.eos_id, pad_token_id=l.pad_id,
                     attn_implementation="eager", tie_word_embeddings=False)
    m = LlamaForCausalLM(hc).eval()
    hsd = {k[len("llama_model."):]: v for k, v in sd.items() if k.startswith("llama_model.")}
    missing, unexpected = m.load_state_dict(hsd, strict=False)
    assert not unexpected and all("rotary" in k or "inv_freq" in k for k in missing), (missing, unexpected)
    return m
