"""Parity on architectures OTHER than the two benchmarked ones: the builder is parameterised on model_channels,
channel_mult, the class count and the image shape (note under the Cityscapes table: the released
256x512 checkpoint may use base_channels 64), so the kernels' tilings, weight packings and statistics layouts must not
have the benchmark shapes baked in.  Checker: the CPU oracle's torch fp32 restatement (oracle/unet_ref.py) on the same
seeded inputs and weights; tolerances are the ones of tests/test_gpu_chain.py.
"""
import numpy as np
import pytest
import torch

from conftest import DINO, UNET_PARAMS
from test_gpu_chain import (BF16_LABEL_FRAC, BF16_X0_MAX, BF16_X0_MEAN, PARITY_MODES, X0_TOL, _onehot, _report,
                            _teacher_forced_check)

pytestmark = pytest.mark.gpu

VARIANTS = {
    # tag: (base_channels, channel_mult, attention at downsample rates, C_img, H, W, K, B)
    "wide64": (64, None, [32, 16, 8], 1, 64, 64, 2, 3),      # image_size 64 -> (1, 2, 3, 4): 64..256 channels, 512-channel concats
    "k5_rgb_32x96": (32, (1, 1, 2), [2, 4], 3, 32, 96, 5, 2),  # attention over 768 and 192 tokens (partial query / key tiles), non-square
    "k19_odd_batch": (32, (1, 2, 2), [1, 4], 3, 64, 64, 19, 5),  # attention on the full-resolution map (4096 tokens), 64-channel first level
    "deep_lidc_b1": (32, (1, 1, 2, 3, 4), [32, 16, 8], 1, 128, 128, 2, 1),
    "lidc_48x80": (32, (1, 1, 2, 3, 4), [32, 16, 8], 1, 48, 80, 2, 2),   # 3x5 maps at the bottom: attention over 15 and 60 tokens
    "head_dim_64": (64, (1, 2), [1, 2], 1, 32, 32, 3, 2),                # num_head_channels = 64
}
VARIANTS.update({
    "size512_mult": (64, (0.5, 1, 1, 2, 2, 4, 4), [32, 16, 8], 1, 128, 128, 2, 1),  # seven levels, 2x2 maps at the bottom
    "cs_vitb_768": (32, (1, 1, 2, 2, 4, 4), [32, 16, 8], 3, 64, 128, 20, 2),         # a 768-channel feature condition (dino_vitb8)
    # base_channels 64 with the DINO concat: 128 + 384 channels, groups of 16 -- no GroupNorm group straddles the boundary, so the
    # per-step half of the folded block takes no feature channels at all
    "cs_base64_dino": (64, (1, 1, 2, 2), [8], 3, 64, 128, 20, 2),
})
EXTRA = {"head_dim_64": dict(num_head_channels=64)}
FCE = {"cs_vitb_768": dict(DINO, model="dino_vitb8", channels=768), "cs_base64_dino": dict(DINO)}
ONLY = {}


def _variant(tag):
    from ccdm_b200 import models
    from ccdm_b200.synthetic import fill_synthetic_, synthetic_inputs
    base, mult, att, C_img, H, W, K, B = VARIANTS[tag]
    p = dict(UNET_PARAMS, base_channels=base, channel_mult=mult, attention_resolutions=att, **EXTRA.get(tag, {}))
    m = models.build_model(50, "cosine", {"s": 0.008}, [(C_img, H, W), (K, H, W)], (C_img, H, W), "unet_openai", p,
                           "datasets.lidc", "majority", FCE.get(tag)).eval()
    fill_synthetic_(m.unet, 3)
    image, feat, labels = synthetic_inputs(B, C_img, H, W, K, FCE[tag]["channels"] if tag in FCE else 0)
    m.test_feat = feat
    return m.cuda(), image, labels, (B, C_img, H, W, K)


@pytest.mark.parametrize("prec", PARITY_MODES + ["bf16"])
@pytest.mark.parametrize("tag", list(VARIANTS))
def test_unet_variant_vs_oracle(cuda_device, tag, prec):
    from oracle import unet_ref
    if prec not in ONLY.get(tag, (prec,)):
        pytest.skip(f"{tag}: not implemented in '{prec}' (refused at plan time, tests/test_host_logic.py)")
    m, image, labels, (B, C_img, H, W, K) = _variant(tag)
    m.unet.precision = prec
    t = torch.full((B,), 23.0)
    feat = m.test_feat
    ref = unet_ref.unet_forward({k: v.cpu() for k, v in m.unet.state_dict().items()}, _onehot(labels, K), image, feat, t,
                                head_channels=EXTRA.get(tag, {}).get("num_head_channels", 32),
                                feature_condition_idx=FCE[tag]["target_layer"] if tag in FCE else None)
    got = m.unet(_onehot(labels, K).cuda(), image.cuda(), feat.cuda() if feat is not None else None, t.cuda())["diffusion_out"].cpu()
    err = (got - ref).abs()
    _report(f"variant_{prec}_{tag}", max_abs_err=err.max(), mean_abs_err=err.mean())
    if prec == "bf16":
        assert float(err.max()) <= BF16_X0_MAX and float(err.mean()) <= BF16_X0_MEAN, (float(err.max()), float(err.mean()))
    else:
        assert float(err.max()) <= X0_TOL, float(err.max())
    if prec != "fp32":
        from ccdm_b200 import _lib
        prog = m.unet.engine(prec).program(B, H, W, 1)
        n_conv = sum(1 for o in prog._op_dicts if o["kind"] == _lib.OP_CONV)
        assert prog.n_tc == n_conv and not prog.off_tc, "a conv fell off the tensor-core kernel"


@pytest.mark.parametrize("tag", ["wide64", "k19_odd_batch"])
def test_chain_variant_teacher_forced_vs_oracle(cuda_device, tag):
    from ccdm_b200 import _lib
    from ccdm_b200.models.diffusion_denoising import reverse_t_values
    m, image, labels, (B, C_img, H, W, K) = _variant(tag)
    ts = reverse_t_values(50, 10000 + 4)
    al, ca = m._schedule_host()
    for prec in ("exact", "bf16"):
        record = []
        m.unet.engine(prec).run_chain(_onehot(labels, K).cuda(), image.cuda(), None, ts, al, ca, _lib.DRAW_MAJORITY,
                                      noise="philox", seed=5, record=record)
        stats = _teacher_forced_check(m, image, None, None, K, 50, record)
        _report(f"variant_teacher_forced_{prec}_{tag}", **stats)
        if prec == "exact":
            assert stats["max_dx0"] <= X0_TOL and stats["max_dlogp"] <= 1e-3
            assert stats["mismatch_outside_margin"] == 0
            assert stats["mismatch_total"] <= 1e-4 * stats["pixels"] + 2
        else:
            assert stats["max_dx0"] <= BF16_X0_MAX
            assert stats["mismatch_total"] / stats["pixels"] <= BF16_LABEL_FRAC
