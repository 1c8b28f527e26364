"""LDM conditional CT generator (stage 2 of GuideGen), drop-in for latentdiffusion/ldm/* on the sampler path."""
from .ddim import DDIMSampler  # noqa: F401
from .ddpm import DiffusionWrapper, LatentDiffusion  # noqa: F401
from .openaimodel import UNetModel  # noqa: F401
from .plms import PLMSSampler  # noqa: F401
