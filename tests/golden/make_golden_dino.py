"""Golden fixtures of the DINO ViT condition encoder, produced by the reference's OWN extractor code.

Run here (the build container), never on the GPU box:

    python tests/golden/make_golden_dino.py

``ddpm/models/dino.py`` (ViTExtractor) is imported unmodified; its two absent imports are handled as follows: ``timm`` (only
used for non-DINO checkpoints) is stubbed with an empty module, and the ViT it would download with
``torch.hub.load('facebookresearch/dino:main', ...)`` is handed to its documented ``model=`` argument instead -- the
restatement of that hub model in ``oracle/dino_ref.py`` (pinned separately against transformers.ViTModel by
tests/test_oracle_dino.py), filled with the repo's deterministic synthetic weights.  So the hook mechanics, facet selection,
cls removal, channel interleave, patch-grid reshape, bilinear resize and the stride patch with its position-embedding
interpolation are the reference's code; the transformer blocks are the published architecture.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "ccdm-stochastic-segmentation_b200"))
sys.path.insert(0, HERE)
sys.path.insert(0, "/root/reference/ddpm/models")
_stub = "timm" not in sys.modules
if _stub:
    sys.modules["timm"] = types.ModuleType("timm")
import dino as ref_dino  # noqa: E402  (the reference's ddpm/models/dino.py)
if _stub:
    del sys.modules["timm"]

from ccdm_b200.synthetic import fill_synthetic_  # noqa: E402
from oracle import dino_ref  # noqa: E402

from dino_cases import CASES, image  # noqa: E402


def run_reference(model_type, stride, batch, layers, resize_shape):
    vit = fill_synthetic_(dino_ref.build(model_type), 0).eval()
    ext = ref_dino.ViTExtractor(model_type, stride, model=vit, device="cpu")
    with torch.no_grad():
        return ext.extract_descriptors(batch, layers, resize_shape=resize_shape)


def main():
    for tag, (mt, stride, B, H, W, layers, rs) in CASES.items():
        x = image(B, H, W, 77)
        out = run_reference(mt, stride, x, layers, rs).numpy()
        arrays = dict(shape=np.array(out.shape))
        if out.size <= 200000:
            arrays["desc"] = out
        else:  # benchmark size: a window across every channel, per-channel means and a strided sample
            h0, w0 = min(12, out.shape[2] - 8), min(28, out.shape[3] - 8)
            arrays["win0"] = np.array([h0, w0])
            arrays["window"] = out[:, :, h0:h0 + 8, w0:w0 + 8]
            arrays["chan_mean"] = out.astype(np.float64).mean(axis=(2, 3))
            arrays["strided"] = out[:, ::7, ::3, ::5]
        path = os.path.join(HERE, tag + ".npz")
        np.savez_compressed(path, **arrays)
        print(f"{tag}: {tuple(out.shape)} |x|max {np.abs(out).max():.3f} std {out.std():.3f}  {os.path.getsize(path) / 1024:.1f} KiB")


if __name__ == "__main__":
    main()
