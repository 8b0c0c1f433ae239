from .tsformer import TSFormer

__all__ = ["TSFormer"]
