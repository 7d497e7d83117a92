"""Deterministic synthetic weights and inputs (benchmarks and tests).

A freshly constructed reference model is degenerate: the second conv of every
ResBlock, every attention ``proj_out`` and the output conv are zero-initialised
(/root/reference/ddpm/models/unet_openai/unet.py:216-218,300,705), so it predicts
exactly 1/K everywhere and any parity check passes vacuously (SURVEY.md section 7,
hard part 9).  There is no network for checkpoints, so benchmarks and tests fill
every tensor of a UNet ``state_dict`` from a recipe that depends only on
(seed, key name, shape) -- identical for the reference model, the oracle and this
package, independent of construction order and torch's global generator.
"""
import zlib

import torch


def synthetic_tensor(key: str, shape, seed: int = 0) -> torch.Tensor:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    shape = tuple(shape)
    r = torch.randn(shape, generator=g, dtype=torch.float32)
    if key.endswith("pos_embed") or key.endswith("cls_token"):  # ViT position / class embeddings (DINO encoder)
        return 0.2 * r
    if len(shape) == 1:
        if key.endswith(".weight"):  # GroupNorm gains are the only 1-D weights
            return 1.0 + 0.1 * r
        return 0.1 * r  # biases (conv, linear, GroupNorm)
    fan_in = 1
    for s in shape[1:]:
        fan_in *= s
    return r / float(fan_in) ** 0.5


def synthetic_state_dict(shapes: dict, seed: int = 0) -> dict:
    """``shapes``: {key: shape}.  Returns {key: fp32 CPU tensor}."""
    return {k: synthetic_tensor(k, s, seed) for k, s in shapes.items()}


def fill_synthetic_(module: torch.nn.Module, seed: int = 0) -> torch.nn.Module:
    """In-place fill of every parameter of ``module`` (a UNet with reference key names)."""
    sd = module.state_dict()
    new = {k: synthetic_tensor(k, v.shape, seed).to(v.dtype) for k, v in sd.items()}
    module.load_state_dict(new)
    return module


def synthetic_inputs(B, C_img, H, W, K, feat_channels=0, seed=1234):
    """image ~ N(0,1), optional feature condition ~ N(0,1) at H/8 x W/8, uniform labels x_T."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    image = torch.randn((B, C_img, H, W), generator=g)
    feat = torch.randn((B, feat_channels, H // 8, W // 8), generator=g) if feat_channels else None
    labels = torch.randint(0, K, (B, H, W), generator=g, dtype=torch.uint8)
    return image, feat, labels
