import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "ccdm-stochastic-segmentation_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")
if GOLDEN not in sys.path:
    sys.path.insert(0, GOLDEN)  # dino_cases.py: the case table shared with the fixture generator
REFERENCE = "/root/reference/ddpm"

UNET_PARAMS = dict(base_channels=32, channel_mult=None, attention_resolutions=[32, 16, 8], num_heads=1,
                   num_head_channels=32, softmax_output=True)
DINO = dict(type="dino", model="dino_vits8", channels=384, conditioning="concat_pixels_concat_features",
            output_stride=8, scale="single", train=False, source_layer=11, target_layer=10)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def build_ours(T, C_img, H, W, K, step_T_sample="majority", fce=None, channel_mult=None, seed=0):
    """Our DenoisingModel with the deterministic synthetic weights (CPU tensors)."""
    from ccdm_b200 import models
    from ccdm_b200.synthetic import fill_synthetic_
    p = dict(UNET_PARAMS)
    p["channel_mult"] = channel_mult
    m = models.build_model(T, "cosine", {"s": 0.008}, [(C_img, H, W), (K, H, W)], (C_img, H, W), "unet_openai", p,
                           "datasets.lidc" if K == 2 else "datasets.cityscapes", step_T_sample, fce).eval()
    fill_synthetic_(m.unet, seed)
    return m


GOLDEN_CASES = {
    # tag: (T, B, C_img, H, W, K, fce, channel_mult, t_probe, steps)  -- mirrors tests/golden/make_golden.py
    "lidc64": (250, 2, 1, 64, 64, 2, None, None, (250, 37, 1), 7),
    "lidc128": (250, 1, 1, 128, 128, 2, None, None, (250, 1), 5),
    "cs64x128": (250, 1, 3, 64, 128, 20, DINO, (1, 1, 2, 2, 4, 4), (250, 100), 4),
}


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda", 0)
