"""Drop-in architecture package: the two classes the reference's configs import (``from .step_arch import STEP`` in
STEP_<DATASET>.py, ``TSFormer`` in TSFormer_<DATASET>.py), backed by the step_b200 CUDA library."""
from .step import STEP
from .tsformer import TSFormer

__all__ = ["STEP", "TSFormer"]
