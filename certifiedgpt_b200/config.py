"""Shapes of the MiniGPT-4 towers on the smoothing path (defaults = the reference's models)."""
from dataclasses import dataclass, field


@dataclass
class VitConfig:
    """EVA ViT-g/14, create_eva_vit_g (eva_vit.py:425-437)."""
    img_size: int = 224
    patch: int = 14
    dim: int = 1408
    depth: int = 39
    heads: int = 16
    mlp: int = 6144          # int(1408 * 4.3637)
    eps: float = 1e-6

    @property
    def grid(self):
        return self.img_size // self.patch

    @property
    def tokens(self):
        return self.grid * self.grid + 1

    @property
    def head_dim(self):
        return self.dim // self.heads


@dataclass
class QFormerConfig:
    """BLIP-2 Q-Former: bert-base with cross-attention every 2nd layer (minigpt4.py:90-119)."""
    hidden: int = 768
    layers: int = 12
    heads: int = 12
    inter: int = 3072
    n_query: int = 32
    cross_freq: int = 2
    eps: float = 1e-12

    @property
    def head_dim(self):
        return self.hidden // self.heads

    def cross_layers(self):
        return [i for i in range(self.layers) if i % self.cross_freq == 0]


@dataclass
class LlmConfig:
    """Llama-2-7B / Vicuna (HF LlamaConfig)."""
    hidden: int = 4096
    layers: int = 32
    heads: int = 32
    inter: int = 11008
    vocab: int = 32000
    rms_eps: float = 1e-5
    rope_theta: float = 10000.0
    eos_id: int = 2
    pad_id: int = 0

    @property
    def head_dim(self):
        return self.hidden // self.heads


@dataclass
class ModelConfig:
    vit: VitConfig = field(default_factory=VitConfig)
    qf: QFormerConfig = field(default_factory=QFormerConfig)
    llm: LlmConfig = field(default_factory=LlmConfig)
    ln_vision_eps: float = 1e-5   # nn.LayerNorm default (base_model.py:281-287)

    @staticmethod
    def tiny():
        """Small shapes every kernel supports; used by the parity tests."""
        return ModelConfig(
            vit=VitConfig(img_size=56, dim=64, depth=2, heads=4, mlp=128),
            qf=QFormerConfig(hidden=64, layers=2, heads=4, inter=128, n_query=8),
            llm=LlmConfig(hidden=128, layers=2, heads=4, inter=256, vocab=96),
        )

    @staticmethod
    def full(img_size=224):
        return ModelConfig(vit=VitConfig(img_size=img_size))
