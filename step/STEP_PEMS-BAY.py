"""STEP(PEMS-BAY) configuration - same file name and CFG layout as the reference's step/STEP_PEMS-BAY.py."""
from .configs import step_config

CFG = step_config("PEMS-BAY")
