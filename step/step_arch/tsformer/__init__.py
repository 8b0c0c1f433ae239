"""TSFormer (stage 1 of STEP) on the step_b200 kernels; also exposes the mask generator used in pre-train mode."""
from .tsformer import MaskGenerator, TSFormer

__all__ = ["TSFormer", "MaskGenerator"]
