"""B200-native (sm_100a) drop-in for the hot path of yangkunyi/sam2-video-training:
SAM2 MemoryAttention (RoPE self/cross attention, forward + backward) and the per-frame
multi-object mask losses.  Host side mirrors the reference's module interfaces; compute goes
through the C ABI of libsam2b200.so (include/sam2_b200.h).  No CPU fallback."""
from . import _lib  # noqa: F401

__version__ = "0.1.0"
