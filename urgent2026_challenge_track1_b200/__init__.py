"""urgent2026_challenge_track1_b200 — B200-native (sm_100a) implementation of the BSRNN / BSRNN-FlowSE
speech-enhancement hot path of urgent-challenge/urgent2026_challenge_track1, behind the reference's own model API
(class names, constructor arguments, forward signatures and state_dict layout).

    from urgent2026_challenge_track1_b200 import BSRNN_SE          # baseline_code/models/bsrnn.py:9
    from urgent2026_challenge_track1_b200.bsrnn_flowse import BSRNN # baseline_code/models/bsrnn_flowse.py:171

All arithmetic runs in hand-written CUDA kernels behind the C ABI in include/bsrnn_b200.h; there is no CPU or
PyTorch fallback (a missing extension raises NativeLibraryError).
"""
from ._lib import NativeLibraryError, LIB_PATH  # noqa: F401
from .bsrnn import BSRNN_SE  # noqa: F401

__all__ = ["BSRNN_SE", "NativeLibraryError", "LIB_PATH"]
