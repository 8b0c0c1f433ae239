"""TSFormer(PEMS03) pre-training configuration - same file name and CFG layout as the reference's step/TSFormer_PEMS03.py."""
from .configs import tsformer_config

CFG = tsformer_config("PEMS03")
