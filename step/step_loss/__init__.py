from .step_loss import step_loss

__all__ = ["step_loss"]
