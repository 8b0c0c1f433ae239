"""Runner glue with the BasicTS contract (host tuple in, 4-tuple out, loss splat) - see step_runner.py."""
from .step_runner import STEPRunner
from .tsformer_runner import TSFormerRunner

__all__ = ["STEPRunner", "TSFormerRunner"]
