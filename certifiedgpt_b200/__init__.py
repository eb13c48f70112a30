"""certifiedgpt_b200: B200-native Monte-Carlo randomized-smoothing hot path of CertifiedGPT.

Public surface (mirrors the reference):
    from certifiedgpt_b200.randomized_smoothing.smoothing import Smooth
"""
__version__ = "0.1.0"
