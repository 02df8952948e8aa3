"""ref: recformer/models.py — same module path, names re-exported from recformer_b200.models."""
from recformer_b200.models import *  # noqa: F401,F403
from recformer_b200 import models as _impl

__all__ = [n for n in dir(_impl) if n.startswith("Recformer") or n in ("Similarity",)]
