"""Feature-condition encoder: drop-in for the reference's ``ddpm/models/condition_encoder.py``.

``DinoViT(name, train_encoder, conditioning, stride, resize_shape, layers)`` (:25-46) and ``_build_feature_cond_encoder(params)``
(:55-81) with the same arguments; ``forward(x)`` returns the 'key' descriptors of ViT block ``layers`` as
``[B, channels, H // stride, W // stride]`` fp32 -- the ``feature_condition`` the UNet concatenates (unet.py:770-788).
Inference only (the sampler's path): ``train_encoder=True`` raises.  No ignite here, so ``.to(idist.device())`` becomes
``.to('cuda')`` and the DDP / DataParallel wrappers of the training path are not applied.
"""
import logging
from typing import Union

import torch
from torch import nn

from .dino import ViTExtractor

LOGGER = logging.getLogger(__name__)


class ConditionEncoder(nn.Module):
    def __init__(self):
        super().__init__()


class DinoViT(ConditionEncoder):
    def __init__(self, name: str, train_encoder: bool, conditioning: str, stride: int = 8, resize_shape: Union[tuple, None] = None,
                 layers: Union[list, int] = 11):
        super().__init__()
        if train_encoder:
            raise NotImplementedError("train_encoder=True: the B200 encoder is inference-only (the sampler's path)")
        self.extractor = ViTExtractor(name, stride)
        for param in self.parameters():
            param.requires_grad = False
        self.stride = stride
        self.conditioning = conditioning
        self.layers = layers
        self.resize_shape = resize_shape

    def forward(self, x: torch.Tensor) -> Union[torch.Tensor, list]:
        return self.extractor.extract_descriptors(x, self.layers, resize_shape=self.resize_shape)


def create_cond_fis_fn_default(params):
    mean = torch.tensor([0.485, 0.456, 0.406])
    std = torch.tensor([0.229, 0.224, 0.225])

    def denorm(x):
        return x * std.to(x.device)[:, None, None] + mean.to(x.device)[:, None, None]

    return (lambda x: x / 2 + 0.5) if params["dataset_file"] in ["datasets.lidc", "datasets.lidc_orig"] else denorm


def _build_feature_cond_encoder(params: dict):
    fce_params = params["feature_cond_encoder"]
    if "dino" in fce_params["type"]:
        feature_cond_encoder = DinoViT(fce_params["model"], fce_params["train"], fce_params["conditioning"],
                                       stride=fce_params["output_stride"]).to("cuda")
        LOGGER.info(f"Feature Condition encoder {fce_params} parameters: {sum(p.numel() for p in feature_cond_encoder.parameters())}")
        cond_vis_fn = lambda x: x * torch.tensor([0.229, 0.224, 0.225], device=x.device)[:, None, None] \
            + torch.tensor([0.485, 0.456, 0.406], device=x.device)[:, None, None]  # noqa: E731
    else:
        feature_cond_encoder = None
        cond_vis_fn = create_cond_fis_fn_default(params)
        LOGGER.info("No Feature Condition encoder in use.")
    return feature_cond_encoder, cond_vis_fn
