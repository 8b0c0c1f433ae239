"""Training objective of STEP: masked MAE of the forecast + weighted BCE between theta and the kNN prior graph."""
from .step_loss import step_loss

__all__ = ["step_loss"]
