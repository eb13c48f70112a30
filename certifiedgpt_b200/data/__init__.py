"""Data formats on either side of the smoothing path: the VQAv2 loader and the BLIP-2 processors (host side)."""
