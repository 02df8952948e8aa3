"""recformer_b200 — B200-native drop-in for the Recformer encoder + scoring hot path.

Mirrors the reference `recformer` package's public names (ref: recformer/__init__.py:1-3) for
the hot path: RecformerConfig, RecformerModel, RecformerForSeqRec, RecformerForPretraining,
RecformerForFraudDetection (+ FocalLoss), RecformerTokenizer."""
__all__ = ["RecformerConfig", "RecformerModel", "RecformerForSeqRec", "RecformerForPretraining",
           "RecformerForFraudDetection", "FocalLoss", "RecformerTokenizer", "Ranker", "encode_all_items"]


def __getattr__(name):
    if name in ("RecformerConfig", "RecformerModel", "RecformerForSeqRec", "RecformerForPretraining", "Similarity",
                "RecformerForFraudDetection", "FocalLoss"):
        from . import models
        return getattr(models, name)
    if name == "RecformerTokenizer":
        from .tokenization import RecformerTokenizer
        return RecformerTokenizer
    if name in ("Ranker", "TopKRanker"):
        from . import metrics
        return getattr(metrics, name)
    if name == "encode_all_items":
        from .items import encode_all_items
        return encode_all_items
    raise AttributeError(name)
