"""certifiedgpt_b200: B200-native Monte-Carlo randomized-smoothing hot path of CertifiedGPT.

Public surface (mirrors the reference):
    from certifiedgpt_b200.randomized_smoothing.smoothing import Smooth      # certify / predict, reference API
    from certifiedgpt_b200.native import NativeMiniGPT4Engine                # MiniGPT-4 classifier behind a libcgpt handle
    from certifiedgpt_b200.engine import MiniGPT4Engine                      # the same kernels driven from Python
    from certifiedgpt_b200.agents import setup_agent                         # certify / predict / fine-tune agents
    from certifiedgpt_b200.train import LlamaProjTrainer                     # noise-augmented fine-tune step
    from certifiedgpt_b200.attack import ClipVisionEngine, BlackBoxAttack    # attack inner loop
    from certifiedgpt_b200.data import vqav2                                 # VQAv2 loader, BLIP-2 processors
Everything computes through certifiedgpt_b200/lib/libcgpt.so (include/cgpt.h); there is no CPU fallback.
"""
__version__ = "0.2.0"
