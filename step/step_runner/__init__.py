from .step_runner import STEPRunner

__all__ = ["STEPRunner"]
