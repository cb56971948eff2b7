"""slnlp_b200: B200-native (sm_100a) hot path of sign-language-nlp.

Host side of the C ABI in include/slnlp_b200.h.  Importing this package loads the
CUDA library; there is no CPU or PyTorch-eager fallback for the model math.
"""
from . import _lib  # noqa: F401  (fails loudly when the .so is missing)

__all__ = ["_lib"]
