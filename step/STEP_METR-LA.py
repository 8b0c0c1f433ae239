"""STEP(METR-LA) configuration - same file name and CFG layout as the reference's step/STEP_METR-LA.py."""
from .configs import step_config

CFG = step_config("METR-LA")
