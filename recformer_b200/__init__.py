"""recformer_b200 — B200-native drop-in for the Recformer encoder + scoring hot path.

Mirrors the reference `recformer` package's public names (ref: recformer/__init__.py:1-3) for
the hot path: RecformerConfig, RecformerModel, RecformerForSeqRec, RecformerTokenizer."""
__all__ = ["RecformerConfig", "RecformerModel", "RecformerForSeqRec", "RecformerTokenizer", "Ranker"]


def __getattr__(name):
    if name in ("RecformerConfig", "RecformerModel", "RecformerForSeqRec", "Similarity"):
        from . import models
        return getattr(models, name)
    if name == "RecformerTokenizer":
        from .tokenization import RecformerTokenizer
        return RecformerTokenizer
    if name in ("Ranker", "TopKRanker"):
        from . import metrics
        return getattr(metrics, name)
    raise AttributeError(name)
