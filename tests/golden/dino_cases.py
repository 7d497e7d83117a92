"""Case table of the DINO encoder fixtures (shared by make_golden_dino.py and the tests; imports nothing of the reference)."""
import torch

# tag: (model_type, stride, B, H, W, layers, resize_shape)
CASES = {
    "dino_s8_64x128": ("dino_vits8", 8, 2, 64, 128, 11, None),          # the Cityscapes-shaped small fixture (cs64x128)
    "dino_s8_224": ("dino_vits8", 8, 1, 224, 224, 11, None),            # native grid: pos_embed used as stored
    "dino_s8_stride4_72x104": ("dino_vits8", 4, 1, 72, 104, 11, None),  # stride patch: overlapping patches, 17x25 grid resized to 18x26
    "dino_s8_layer5_resize": ("dino_vits8", 8, 1, 96, 64, 5, (20, 12)), # another layer, explicit resize_shape
    "dino_s8_256x512": ("dino_vits8", 8, 1, 256, 512, 11, None),        # the benchmark size (2049 tokens)
}


def image(B, H, W, seed):
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return torch.randn((B, 3, H, W), generator=g)


