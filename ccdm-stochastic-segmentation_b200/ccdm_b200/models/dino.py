"""ViTExtractor: the reference's descriptor extractor (ddpm/models/dino.py:15-309) over the CUDA ViT engine.

Mirrors what the sampler's feature conditioning uses of it: the constructor ``ViTExtractor(model_type, stride, model, device)``
(:27-56), ``create_model`` (:58-82), ``patch_vit_resolution`` (:119-139) and ``extract_descriptors`` (:279-322) for the 'key',
facet.  The ViT is not a stack of torch layers: ``VisionTransformer`` below only HOLDS the parameters under the key names of
the hub model the reference loads (``torch.hub.load('facebookresearch/dino:main', model_type)``: ``cls_token``, ``pos_embed``,
``patch_embed.proj.*``, ``blocks.N.{norm1,attn.qkv,attn.proj,norm2,mlp.fc1,mlp.fc2}.*``, ``norm.*``), so a state_dict of that
model -- or a checkpoint's ``feature_cond_encoder`` entry (eval_cdm.py:136-142) -- loads strictly; ``forward`` raises, the
arithmetic lives in ``ccdm_b200.vit_engine`` / ``csrc/vit.cu``.  There is no network here: ``create_model`` builds the
architecture with the hub code's initialisation and loads weights from ``$CCDM_DINO_WEIGHTS`` (a ``torch.save``d state_dict
of the hub model) when that is set.  Saliency maps, log-binned descriptors, the 'token' / 'attn' / 'query' / 'value' facets and
image-file preprocessing are not on the sampler's path and raise ``NotImplementedError``.
"""
import os
from typing import List, Tuple, Union

import torch
from torch import nn

# dino.py:60-76 model_type -> (patch, embed_dim, depth, heads): vit_small / vit_base of vision_transformer.py
ARCHS = {"dino_vits8": (8, 384, 12, 6), "dino_vits16": (16, 384, 12, 6), "dino_vitb8": (8, 768, 12, 12), "dino_vitb16": (16, 768, 12, 12)}


class _Holder(nn.Module):
    def forward(self, *a, **k):
        raise NotImplementedError("parameter holder: the ViT runs in ccdm_b200.vit_engine (hand-written sm_100a kernels)")


class _PatchEmbed(_Holder):
    def __init__(self, patch_size, in_chans, embed_dim):
        super().__init__()
        self.patch_size = patch_size
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)


class _Attention(_Holder):
    def __init__(self, dim, num_heads):
        super().__init__()
        self.num_heads = num_heads
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim)


class _Mlp(_Holder):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)


class _Block(_Holder):
    def __init__(self, dim, num_heads):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = _Attention(dim, num_heads)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = _Mlp(dim, 4 * dim)


class VisionTransformer(_Holder):
    """Parameters of the hub ViT under its own key names (vision_transformer.py VisionTransformer.__init__)."""

    def __init__(self, patch_size=8, embed_dim=384, depth=12, num_heads=6, img_size=224, in_chans=3):
        super().__init__()
        self.embed_dim, self.num_heads = embed_dim, num_heads
        self.patch_embed = _PatchEmbed(patch_size, in_chans, embed_dim)
        n = (img_size // patch_size) ** 2
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, n + 1, embed_dim))
        self.blocks = nn.ModuleList([_Block(embed_dim, num_heads) for _ in range(depth)])
        self.norm = nn.LayerNorm(embed_dim, eps=1e-6)
        nn.init.trunc_normal_(self.pos_embed, std=0.02)
        nn.init.trunc_normal_(self.cls_token, std=0.02)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=0.02)
                nn.init.constant_(m.bias, 0)


class ViTExtractor(nn.Module):
    def __init__(self, model_type: str = "dino_vits8", stride: int = 4, model: nn.Module = None, device: str = "cuda"):
        super().__init__()
        self.model_type = model_type
        self.device = device
        self.model = model if model is not None else ViTExtractor.create_model(model_type)
        self.model = ViTExtractor.patch_vit_resolution(self.model, stride=stride)
        self.model.eval()
        self.model.to(self.device)
        self.p = self.model.patch_embed.patch_size
        self.stride = self.model.patch_embed.proj.stride
        self.mean = (0.485, 0.456, 0.406) if "dino" in self.model_type else (0.5, 0.5, 0.5)
        self.std = (0.229, 0.224, 0.225) if "dino" in self.model_type else (0.5, 0.5, 0.5)
        self.load_size = None
        self.num_patches = None
        self._engine = None

    @staticmethod
    def create_model(model_type: str) -> nn.Module:
        if model_type not in ARCHS:
            raise NotImplementedError(f"model_type {model_type}: the B200 encoder builds {sorted(ARCHS)} (timm checkpoints need timm)")
        patch, dim, depth, heads = ARCHS[model_type]
        model = VisionTransformer(patch, dim, depth, heads)
        path = os.environ.get("CCDM_DINO_WEIGHTS")
        if path:
            sd = torch.load(path, map_location="cpu")
            sd = {k: v for k, v in sd.items() if not k.startswith("head.")}
            model.load_state_dict(sd, strict=True)
        return model

    @staticmethod
    def patch_vit_resolution(model: nn.Module, stride: int) -> nn.Module:
        patch_size = model.patch_embed.patch_size
        stride = nn.modules.utils._pair(stride)
        assert all([(patch_size // s_) * s_ == patch_size for s_ in stride]), f"stride {stride} should divide patch_size {patch_size}"
        if stride[0] != stride[1]:
            raise NotImplementedError("the B200 encoder takes one stride for both axes")
        model.patch_embed.proj.stride = stride  # (the position-embedding rule of dino.py:86-117 is applied by the engine)
        return model

    def engine(self):
        from ..vit_engine import ViTEngine
        if self._engine is None or self._engine.device != next(self.model.parameters()).device:
            self._engine = ViTEngine(self.model, int(self.stride[0]))
        return self._engine

    @torch.no_grad()
    def extract_descriptors(self, batch: torch.Tensor, layers: Union[int, list] = 11, facet: str = "key", include_cls: bool = False,
                            resize_shape: Union[tuple, None] = None) -> Union[torch.Tensor, list]:
        assert facet in ["key", "query", "value", "token"], f"{facet} is not a supported facet for descriptors."
        if facet != "key" or include_cls:
            raise NotImplementedError("the B200 encoder extracts the 'key' facet without the cls token (what DinoViT.forward asks for)")
        B, C, H, W = batch.shape
        s = int(self.stride[0])
        self.load_size = (H, W)
        self.num_patches = (1 + (H - self.p) // s, 1 + (W - self.p) // s)
        eng = self.engine()
        if type(layers) == int:
            size = (H // s, W // s) if resize_shape is None else tuple(resize_shape)
            return eng.key_descriptors(batch, [layers], [size])[0]
        elif isinstance(layers, list):
            if B != 1:
                raise ValueError("a list of layers takes one image at a time (dino.py:313 views the descriptors as batch 1)")
            return eng.key_descriptors(batch, list(layers), [tuple(resize_shape) if resize_shape is not None else None] * len(layers))
        raise TypeError("layers must be an int or a list of ints")

    def preprocess(self, *a, **k):
        raise NotImplementedError("image-file preprocessing is outside the sampler's path")

    def extract_saliency_maps(self, *a, **k):
        raise NotImplementedError("saliency maps are outside the sampler's path")
