"""``build_model``: same signature and wiring as the reference (ddpm/models/builder.py:14-51)."""
import logging
from typing import Any, Dict, List, Tuple, Union

import torch

from .diffusion_denoising import DenoisingModel, DiffusionModel
from .unet_openai import create_unet_openai

LOGGER = logging.getLogger(__name__)


def build_model(time_steps: int, schedule: str, schedule_params: Union[dict, None],
                input_shapes: List[Tuple[int, int, int]], cond_encoded_shape, backbone: str,
                backbone_params: Dict[str, Any], dataset_file: str, step_T_sample: str = None,
                feature_cond_encoder: dict = None) -> DenoisingModel:
    img_shape, label_shape = input_shapes
    num_classes = label_shape[0]
    diffusion = DiffusionModel(schedule, time_steps, num_classes, schedule_params=schedule_params)
    if backbone != "unet_openai":
        raise NotImplementedError(f"backbone {backbone}")
    unet = create_unet_openai(image_size=min(img_shape[1], img_shape[2]), in_channels=num_classes + img_shape[0],
                              out_channels=num_classes, num_res_blocks=2, cond_encoded_shape=cond_encoded_shape,
                              feature_cond_encoder=feature_cond_encoder, **backbone_params)
    LOGGER.info("%s trainable params: %d", backbone, sum(map(torch.numel, unet.parameters())))
    return DenoisingModel(diffusion, unet, dataset_file, step_T_sample)
