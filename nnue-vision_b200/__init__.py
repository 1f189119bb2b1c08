"""B200-native NNUE hot path (drop-in for the NNUE model of marict/nnue-vision).

    from nnue_vision_b200 import nnue, serialize, engine, train

`nnue`      NNUE / FeatureTransformer / SimpleClassifier / GridFeatureSet / LossParams (float training path)
`serialize` .nnue v2 writer + quantiser (byte-identical to the reference's serialize.py)
`engine`    NNUEEvaluator: batched bit-exact integer inference
`train`     compute_loss + the data-parallel training step (flat gradient buffer, NCCL all-reduce) + fused optimizer
`evaluate`  evaluate_model / evaluate_compiled_model with the reference's metric names (batched integer engine)

Everything compute-heavy goes through libnnue_b200.so (include/nnue_b200.h); there is no fallback.
"""
from . import _lib  # noqa: F401
from .nnue import NNUE, FeatureTransformer, GridFeatureSet, LossParams, SimpleClassifier  # noqa: F401
from .serialize import serialize_model  # noqa: F401

__all__ = ["NNUE", "FeatureTransformer", "SimpleClassifier", "GridFeatureSet", "LossParams", "serialize_model"]
