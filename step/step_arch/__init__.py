from .tsformer import TSFormer
from .step import STEP

__all__ = ["TSFormer", "STEP"]
