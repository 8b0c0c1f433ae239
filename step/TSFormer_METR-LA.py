"""TSFormer(METR-LA) pre-training configuration - same file name and CFG layout as the reference's step/TSFormer_METR-LA.py."""
from .configs import tsformer_config

CFG = tsformer_config("METR-LA")
