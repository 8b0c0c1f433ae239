"""Cosine-similarity Gram matrix on the B200-native kernel (reference: step/step_arch/similarity.py:6-16)."""
from step_b200 import ops


def batch_cosine_similarity(x, y=None):
    """x: [B, N, D] -> [B, N, N].  Only the self-similarity case (y is x) exists on the STEP path."""
    if y is not None and y is not x:
        raise NotImplementedError("step_b200 implements the self-similarity case used by STEP (x is y)")
    return ops.cosine_gram(x)
