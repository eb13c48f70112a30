"""The two agents the reference leaves as 0-byte files (agents/minigpt4_certify_agent.py,
agents/minigpt4_predict_agent.py), written against the drop-in Smooth, and the noise-augmented fine-tune agent
(agents/minigpt4_finetune_agent.py) on the libcgpt training step.  They keep the reference's
agent protocol (setup_agent / run / finalize, launch.py:105-107) without torch_xla."""
from .minigpt4_certify_agent import MiniGPT4CertifyAgent  # noqa: F401
from .minigpt4_finetune_agent import MiniGPT4FineTuneAgent  # noqa: F401
from .minigpt4_predict_agent import MiniGPT4PredictAgent  # noqa: F401

AGENTS = {"image_text_certify": MiniGPT4CertifyAgent, "image_text_predict": MiniGPT4PredictAgent,
          "image_text_finetune": MiniGPT4FineTuneAgent}


def setup_agent(name, **kwargs):
    """registry.get_agent_class(name).setup_agent(cfg=...) of agents/__init__.py:14-21."""
    assert name in AGENTS, f"Agent {name} not properly registered."
    return AGENTS[name].setup_agent(**kwargs)
