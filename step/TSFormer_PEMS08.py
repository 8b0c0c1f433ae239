"""TSFormer(PEMS08) pre-training configuration - same file name and CFG layout as the reference's step/TSFormer_PEMS08.py."""
from .configs import tsformer_config

CFG = tsformer_config("PEMS08")
