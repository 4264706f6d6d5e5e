"""`from diffusers import DPMSolverMultistepScheduler` -> lidiff_b200.scheduler"""
from lidiff_b200.scheduler import DPMSolverMultistepScheduler  # noqa: F401

__version__ = "0.18.0+lidiff_b200"
