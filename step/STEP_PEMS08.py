"""STEP(PEMS08) configuration - same file name and CFG layout as the reference's step/STEP_PEMS08.py."""
from .configs import step_config

CFG = step_config("PEMS08")
