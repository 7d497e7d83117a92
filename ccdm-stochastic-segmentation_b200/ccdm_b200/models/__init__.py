"""Drop-in for the reference's ``ddpm.models`` package (ddpm/models/__init__.py)."""
from .builder import build_model
from .condition_encoder import DinoViT, _build_feature_cond_encoder
from .diffusion_denoising import DenoisingModel, DiffusionModel
from .one_hot_categorical import OneHotCategoricalBCHW
from .unet_openai import create_unet_openai
from .unet_openai.unet import UNetModel

__all__ = ["build_model", "DenoisingModel", "DiffusionModel", "OneHotCategoricalBCHW", "create_unet_openai", "UNetModel", "DinoViT",
           "_build_feature_cond_encoder"]
