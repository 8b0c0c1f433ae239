"""STEP(PEMS03) configuration - same file name and CFG layout as the reference's step/STEP_PEMS03.py."""
from .configs import step_config

CFG = step_config("PEMS03")
