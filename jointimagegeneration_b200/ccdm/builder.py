"""Drop-in for ccdm/ddpm/models/builder.py::build_model (:14-53)."""
import logging
from typing import Any, Dict, List, Tuple, Union

import torch

from .diffusion_denoising import DenoisingModel, DiffusionModel
from .unet import create_unet_openai

LOGGER = logging.getLogger(__name__)


def build_model(time_steps: int, schedule: str, schedule_params: Union[dict, None], input_shapes: List[Tuple[int, ...]],
                cond_encoded_shape, backbone: str, backbone_params: Dict[str, Any], dataset_file: str,
                step_T_sample: str = None, feature_cond_encoder: dict = None, dims: int = 3) -> DenoisingModel:
    img_shape, label_shape, *_ = input_shapes
    img_channels, num_classes = img_shape[0], label_shape[0]
    diffusion = DiffusionModel(schedule, time_steps, num_classes, schedule_params=schedule_params, dims=dims)
    if backbone != "unet_openai":
        raise NotImplementedError(f"backbone {backbone}")
    model = create_unet_openai(image_size=min(img_shape[1], img_shape[2]), in_channels=num_classes + img_channels,
                               out_channels=num_classes, num_res_blocks=2, cond_encoded_shape=cond_encoded_shape,
                               feature_cond_encoder=feature_cond_encoder, dims=dims, **backbone_params)
    LOGGER.info("%s trainable params: %d", backbone, sum(map(torch.numel, model.parameters())))
    return DenoisingModel(diffusion, model, dataset_file, step_T_sample, dims)
