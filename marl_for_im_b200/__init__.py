"""B200-native batched inventory-management environments (drop-in for the env step of
MarwanMousa/MARL-for-IM).  The CUDA library is loaded lazily on first use and there is
no CPU fallback: creating an env without ``libimx_b200.so`` or without a GPU raises."""
from . import presets  # noqa: F401

__version__ = "0.1.0"
