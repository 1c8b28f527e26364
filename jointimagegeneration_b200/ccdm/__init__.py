"""CCDM volumetric mask sampler (stage 1 of GuideGen), drop-in for ccdm/ddpm/models/*."""
from .builder import build_model  # noqa: F401
from .diffusion_denoising import DenoisingModel, DiffusionModel, cosine_schedule, linear_schedule  # noqa: F401
from .encoder import FrozenBERTEmbedder, PreloadedBERTEncoder  # noqa: F401
from .one_hot_categorical import OneHotCategoricalBCHW  # noqa: F401
from .unet import UNetModel, create_unet_openai  # noqa: F401
