"""`recformer` — the reference's package name, served by the B200-native implementation.

The reference scripts import their model classes with `from recformer import RecformerModel, RecformerForSeqRec,
RecformerTokenizer, RecformerConfig` (ref: recformer/__init__.py:1-3 star-exports models.py + tokenization.py;
finetune.py:12, evaluate_seq.py, cluster.py; finetune_classification.py:16 takes RecformerForFraudDetection).  Putting this repository first on `sys.path` therefore switches an
unmodified `finetune.py` / `evaluate_seq.py` onto the CUDA kernels of `recformer_b200` — nothing here computes.
(`litmodels.LitWrapper`, the Lightning wrapper of lightning_pretrain.py, is control plane and out of scope:
SURVEY.md §2 row 12.)
"""
from recformer_b200.config import RecformerConfig
from recformer_b200.models import (FocalLoss, RecformerForFraudDetection, RecformerForPretraining, RecformerForSeqRec,
                                   RecformerModel, RecformerModelOutput, RecformerPooler, RecformerPretrainingOutput,
                                   Similarity)
from recformer_b200.tokenization import RecformerTokenizer

__all__ = ["RecformerConfig", "RecformerModel", "RecformerForSeqRec", "RecformerForPretraining", "RecformerTokenizer",
           "RecformerForFraudDetection", "FocalLoss", "RecformerPooler", "Similarity", "RecformerModelOutput",
           "RecformerPretrainingOutput"]
