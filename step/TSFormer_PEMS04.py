"""TSFormer(PEMS04) pre-training configuration - same file name and CFG layout as the reference's step/TSFormer_PEMS04.py."""
from .configs import tsformer_config

CFG = tsformer_config("PEMS04")
