from .smoothing import Smooth  # noqa: F401
