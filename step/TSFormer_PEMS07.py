"""TSFormer(PEMS07) pre-training configuration - same file name and CFG layout as the reference's step/TSFormer_PEMS07.py."""
from .configs import tsformer_config

CFG = tsformer_config("PEMS07")
