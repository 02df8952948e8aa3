"""ref: recformer/tokenization.py — same module path, names re-exported from recformer_b200.tokenization."""
from recformer_b200.tokenization import *  # noqa: F401,F403
from recformer_b200 import tokenization as _impl

__all__ = [n for n in dir(_impl) if n.startswith("Recformer") or n in ("Similarity",)]
