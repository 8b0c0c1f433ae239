"""Graph WaveNet backbone of STEP: parameter holder + prologue/epilogue around ``step_b200.ops.GWNetStack``."""
from .model import GraphWaveNet

__all__ = ["GraphWaveNet"]
