"""CPU tests of the DINO condition-encoder oracle (oracle/dino_ref.py) and of the host mirror.

Pinning of the oracle:
  * the extraction logic (hook / key facet / cls removal / channel interleave / resize / stride patch / position-embedding
    interpolation) against fixtures produced by the reference's own ``ViTExtractor`` (tests/golden/make_golden_dino.py), and
    against that class itself when /root/reference is present;
  * the ViT blocks -- a torch.hub dependency that is absent here -- against an independent implementation of the same published
    architecture, ``transformers.ViTModel`` with the weights copied over.
Tolerance: 2e-5 absolute on descriptors of unit scale (CPU fp32 summation order varies with the thread count).
"""
import os
import sys
import types

import numpy as np
import pytest
import torch

from conftest import GOLDEN, golden
from ccdm_b200.synthetic import fill_synthetic_
from oracle import dino_ref

from dino_cases import CASES, image  # noqa: E402

TOL = 2e-5


def _oracle(mt, stride, x, layers, rs):
    vit = fill_synthetic_(dino_ref.build(mt), 0).eval()
    return dino_ref.extract_descriptors(vit, x, layers, stride, rs)


def check_against_golden(out, g):
    assert tuple(out.shape) == tuple(g["shape"])
    if "desc" in g:
        return float(np.abs(out - g["desc"]).max())
    h0, w0 = (int(v) for v in g["win0"])
    e1 = np.abs(out[:, :, h0:h0 + 8, w0:w0 + 8] - g["window"]).max()
    e2 = np.abs(out.astype(np.float64).mean(axis=(2, 3)) - g["chan_mean"]).max()
    e3 = np.abs(out[:, ::7, ::3, ::5] - g["strided"]).max()
    return float(max(e1, e2, e3))


@pytest.mark.parametrize("tag", sorted(CASES))
def test_oracle_reproduces_reference_fixture(tag):
    mt, stride, B, H, W, layers, rs = CASES[tag]
    out = _oracle(mt, stride, image(B, H, W, 77), layers, rs).numpy()
    assert check_against_golden(out, golden(tag + ".npz")) <= TOL


def test_oracle_blocks_agree_with_transformers_vit():
    """Same weights in transformers.ViTModel (an independent implementation of the ViT of Dosovitskiy et al. that DINO's
    vision_transformer.py also implements): final normalised tokens and the key projection of the last block agree."""
    from transformers import ViTConfig, ViTModel
    torch.manual_seed(0)
    vit = fill_synthetic_(dino_ref.build("dino_vits8"), 0).eval()
    cfg = ViTConfig(hidden_size=384, num_hidden_layers=12, num_attention_heads=6, intermediate_size=1536, hidden_act="gelu",
                    layer_norm_eps=1e-6, image_size=224, patch_size=8, num_channels=3, qkv_bias=True, hidden_dropout_prob=0.0,
                    attention_probs_dropout_prob=0.0)
    hf = ViTModel(cfg, add_pooling_layer=False).eval()
    sd = vit.state_dict()
    new = {"embeddings.cls_token": sd["cls_token"], "embeddings.position_embeddings": sd["pos_embed"],
           "embeddings.patch_embeddings.projection.weight": sd["patch_embed.proj.weight"],
           "embeddings.patch_embeddings.projection.bias": sd["patch_embed.proj.bias"],
           "layernorm.weight": sd["norm.weight"], "layernorm.bias": sd["norm.bias"]}
    for i in range(12):
        p, q = f"blocks.{i}.", f"encoder.layer.{i}."
        qw, qb = sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"]
        for j, n in enumerate(("query", "key", "value")):
            new[q + f"attention.attention.{n}.weight"] = qw[384 * j:384 * (j + 1)]
            new[q + f"attention.attention.{n}.bias"] = qb[384 * j:384 * (j + 1)]
        new[q + "attention.output.dense.weight"], new[q + "attention.output.dense.bias"] = sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"]
        new[q + "layernorm_before.weight"], new[q + "layernorm_before.bias"] = sd[p + "norm1.weight"], sd[p + "norm1.bias"]
        new[q + "layernorm_after.weight"], new[q + "layernorm_after.bias"] = sd[p + "norm2.weight"], sd[p + "norm2.bias"]
        new[q + "intermediate.dense.weight"], new[q + "intermediate.dense.bias"] = sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"]
        new[q + "output.dense.weight"], new[q + "output.dense.bias"] = sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"]
    missing, unexpected = hf.load_state_dict(new, strict=False)
    assert not unexpected and not [m for m in missing if "pooler" not in m], (missing, unexpected)
    x = image(1, 224, 224, 5)
    with torch.no_grad():
        ref = hf(pixel_values=x, output_hidden_states=True)
        ours_cls = vit(x)
        assert float((ref.last_hidden_state[:, 0] - ours_cls).abs().max()) <= 5e-5
        lay = hf.encoder.layer[11]
        k_hf = lay.attention.attention.key(lay.layernorm_before(ref.hidden_states[11]))  # [1, T, 384], channel = h * 64 + d
        k = dino_ref.key_facets(vit, x, [11], 8)[11]                                  # [1, heads, T, d]
        assert float((k.permute(0, 2, 1, 3).reshape(1, -1, 384) - k_hf).abs().max()) <= 5e-5


@pytest.mark.skipif(not os.path.exists("/root/reference/ddpm/models/dino.py"), reason="the reference is only present in the build container")
@pytest.mark.parametrize("tag", ["dino_s8_64x128", "dino_s8_stride4_72x104", "dino_s8_layer5_resize"])
def test_oracle_agrees_with_live_reference_extractor(tag):
    stub = "timm" not in sys.modules
    if stub:
        sys.modules["timm"] = types.ModuleType("timm")  # dino.py imports it for non-DINO checkpoints only
    sys.path.insert(0, "/root/reference/ddpm/models")
    try:
        import dino as ref_dino
    finally:
        sys.path.remove("/root/reference/ddpm/models")
        if stub:
            del sys.modules["timm"]
    mt, stride, B, H, W, layers, rs = CASES[tag]
    x = image(B, H, W, 123)  # a seed the fixtures do not use
    vit = fill_synthetic_(dino_ref.build(mt), 3).eval()
    ext = ref_dino.ViTExtractor(mt, stride, model=vit, device="cpu")
    with torch.no_grad():
        want = ext.extract_descriptors(x, layers, resize_shape=rs)
    vit2 = fill_synthetic_(dino_ref.build(mt), 3).eval()  # (the reference patched the first module's stride / method in place)
    got = dino_ref.extract_descriptors(vit2, x, layers, stride, rs)
    assert float((got - want).abs().max()) <= TOL
    # multi-layer branch (dino.py:307-322), batch 1
    want_l = ext.extract_descriptors(x[:1], [2, 5], resize_shape=None)
    got_l = dino_ref.extract_descriptors(vit2, x[:1], [2, 5], stride, None)
    assert all(float((a - b).abs().max()) <= TOL for a, b in zip(got_l, want_l))


def test_host_mirror_has_the_hub_models_state_dict():
    """Our parameter holder exposes exactly the hub ViT's keys and shapes (so its checkpoints load strictly), for every
    model_type the reference's create_model knows of the DINO family."""
    from ccdm_b200.models.dino import ARCHS, VisionTransformer
    for mt, (patch, dim, depth, heads) in ARCHS.items():
        ours = VisionTransformer(patch, dim, depth, heads)
        ref = dino_ref.build(mt)
        a = {k: tuple(v.shape) for k, v in ours.state_dict().items()}
        b = {k: tuple(v.shape) for k, v in ref.state_dict().items()}
        assert a == b, mt
        ours.load_state_dict(ref.state_dict(), strict=True)


def test_host_mirror_errors():
    from ccdm_b200 import _lib
    from ccdm_b200.models.condition_encoder import DinoViT, _build_feature_cond_encoder
    from ccdm_b200.models.dino import ViTExtractor
    ext = ViTExtractor("dino_vits8", 4, device="cpu")
    assert ext.p == 8 and tuple(ext.stride) == (4, 4)
    with pytest.raises(AssertionError, match="should divide patch_size"):
        ViTExtractor("dino_vits8", 3, device="cpu")  # dino.py:131-132
    with pytest.raises(NotImplementedError):
        ViTExtractor("vit_small_patch8_224", 8, device="cpu")  # timm checkpoints (dino.py:65-81): no timm here
    with pytest.raises(AssertionError, match="not a supported facet"):
        ext.extract_descriptors(torch.zeros(1, 3, 32, 32), 11, facet="attn")  # dino.py:290-291
    with pytest.raises(NotImplementedError):
        DinoViT("dino_vits8", True, "concat_pixels_concat_features")
    if not torch.cuda.is_available():
        with pytest.raises((_lib.CcdmError, RuntimeError, AssertionError)):  # no CPU path: loud without a device
            ext.extract_descriptors(torch.zeros(1, 3, 32, 32), 11)
    enc, vis = _build_feature_cond_encoder(dict(feature_cond_encoder=dict(type="none"), dataset_file="datasets.cityscapes"))
    assert enc is None and vis(torch.zeros(3, 2, 2)).shape == (3, 2, 2)
