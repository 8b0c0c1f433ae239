"""TSFormer(PEMS-BAY) pre-training configuration - same file name and CFG layout as the reference's step/TSFormer_PEMS-BAY.py."""
from .configs import tsformer_config

CFG = tsformer_config("PEMS-BAY")
