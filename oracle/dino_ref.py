"""CPU oracle of the DINO ViT condition encoder.  TEST INFRASTRUCTURE ONLY: imported by tests/, the golden
generator and bench.py's CPU legs, never by ccdm_b200.

Two parts:

* ``VisionTransformer`` -- a torch fp32 restatement of the ViT the reference obtains with
  ``torch.hub.load('facebookresearch/dino:main', 'dino_vits8')`` (ddpm/models/dino.py:63), i.e. a THIRD-PARTY dependency that
  is absent from /root/reference: facebookresearch/dino, branch main, ``vision_transformer.py`` (``PatchEmbed``, ``Attention``,
  ``Mlp``, ``Block``, ``VisionTransformer.prepare_tokens / interpolate_pos_encoding / forward``; LayerNorm eps 1e-6, qkv_bias,
  GELU in the erf form, pre-LN residual blocks, cls token first).  Parameter names follow that file, so a state_dict of the hub
  model loads.  Pinned by tests/test_oracle_dino.py against (a) an independent implementation of the same published
  architecture (transformers.ViTModel, weights copied) and (b) the reference's own ViTExtractor (dino.py) driving this module
  through its ``model=`` argument, including its stride patch (``patch_vit_resolution``, dino.py:119-139).
* ``extract_descriptors`` -- restatement of ViTExtractor._extract_features + extract_descriptors for the 'key' facet
  (dino.py:172-176, 211-229, 279-309) and of ``_fix_pos_enc`` (dino.py:86-117), which equals the hub model's own
  interpolate_pos_encoding when stride == patch size.
"""
import math

import torch
import torch.nn as nn

# dino.py:60-76 model_type -> (patch, embed_dim, depth, heads) of vision_transformer.py's vit_small / vit_base
ARCHS = {"dino_vits8": (8, 384, 12, 6), "dino_vits16": (16, 384, 12, 6), "dino_vitb8": (8, 768, 12, 12), "dino_vitb16": (16, 768, 12, 12)}


class Attention(nn.Module):
    def __init__(self, dim, num_heads):
        super().__init__()
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.attn_drop = nn.Dropout(0.0)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(0.0)

    def forward(self, x):
        B, N, C = x.shape
        qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, C // self.num_heads).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0], qkv[1], qkv[2]
        attn = (q @ k.transpose(-2, -1)) * self.scale
        attn = self.attn_drop(attn.softmax(dim=-1))
        x = (attn @ v).transpose(1, 2).reshape(B, N, C)
        return self.proj_drop(self.proj(x)), attn


class Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(hidden, dim)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class Block(nn.Module):
    def __init__(self, dim, num_heads, mlp_ratio=4.0):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = Attention(dim, num_heads)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))

    def forward(self, x):
        y, _ = self.attn(self.norm1(x))
        x = x + y
        return x + self.mlp(self.norm2(x))


class PatchEmbed(nn.Module):
    def __init__(self, img_size, patch_size, in_chans, embed_dim):
        super().__init__()
        self.patch_size = patch_size
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)

    def forward(self, x):
        return self.proj(x).flatten(2).transpose(1, 2)


class VisionTransformer(nn.Module):
    def __init__(self, patch_size=8, embed_dim=384, depth=12, num_heads=6, img_size=224, in_chans=3):
        super().__init__()
        self.embed_dim = embed_dim
        self.patch_embed = PatchEmbed(img_size, patch_size, in_chans, embed_dim)
        n = (img_size // patch_size) ** 2
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, n + 1, embed_dim))
        self.pos_drop = nn.Dropout(0.0)
        self.blocks = nn.ModuleList([Block(embed_dim, num_heads) for _ in range(depth)])
        self.norm = nn.LayerNorm(embed_dim, eps=1e-6)
        self.head = nn.Identity()

    def interpolate_pos_encoding(self, x, w, h):
        # vision_transformer.py VisionTransformer.interpolate_pos_encoding (stride == patch size)
        npatch = x.shape[1] - 1
        N = self.pos_embed.shape[1] - 1
        if npatch == N and w == h:
            return self.pos_embed
        class_pos_embed = self.pos_embed[:, 0]
        patch_pos_embed = self.pos_embed[:, 1:]
        dim = x.shape[-1]
        w0 = w // self.patch_embed.patch_size
        h0 = h // self.patch_embed.patch_size
        w0, h0 = w0 + 0.1, h0 + 0.1
        patch_pos_embed = nn.functional.interpolate(
            patch_pos_embed.reshape(1, int(math.sqrt(N)), int(math.sqrt(N)), dim).permute(0, 3, 1, 2),
            scale_factor=(w0 / math.sqrt(N), h0 / math.sqrt(N)), mode="bicubic")
        assert int(w0) == patch_pos_embed.shape[-2] and int(h0) == patch_pos_embed.shape[-1]
        patch_pos_embed = patch_pos_embed.permute(0, 2, 3, 1).view(1, -1, dim)
        return torch.cat((class_pos_embed.unsqueeze(0), patch_pos_embed), dim=1)

    def prepare_tokens(self, x):
        B, nc, w, h = x.shape
        x = self.patch_embed(x)
        cls_tokens = self.cls_token.expand(B, -1, -1)
        x = torch.cat((cls_tokens, x), dim=1)
        x = x + self.interpolate_pos_encoding(x, w, h)
        return self.pos_drop(x)

    def forward(self, x):
        x = self.prepare_tokens(x)
        for blk in self.blocks:
            x = blk(x)
        x = self.norm(x)
        return x[:, 0]


def build(model_type="dino_vits8"):
    patch, dim, depth, heads = ARCHS[model_type]
    return VisionTransformer(patch, dim, depth, heads)


def fix_pos_enc(model, n_tokens, w, h, patch_size, stride_hw):
    """dino.py:86-117 ``_fix_pos_enc`` as a function of the model (w, h = image height, width: the hub code's naming)."""
    npatch = n_tokens - 1
    N = model.pos_embed.shape[1] - 1
    if npatch == N and w == h:
        return model.pos_embed
    class_pos_embed = model.pos_embed[:, 0]
    patch_pos_embed = model.pos_embed[:, 1:]
    dim = model.pos_embed.shape[-1]
    w0 = 1 + (w - patch_size) // stride_hw[1]
    h0 = 1 + (h - patch_size) // stride_hw[0]
    assert w0 * h0 == npatch
    w0, h0 = w0 + 0.1, h0 + 0.1
    patch_pos_embed = nn.functional.interpolate(
        patch_pos_embed.reshape(1, int(math.sqrt(N)), int(math.sqrt(N)), dim).permute(0, 3, 1, 2),
        scale_factor=(w0 / math.sqrt(N), h0 / math.sqrt(N)), mode="bicubic", align_corners=False, recompute_scale_factor=False)
    assert int(w0) == patch_pos_embed.shape[-2] and int(h0) == patch_pos_embed.shape[-1]
    patch_pos_embed = patch_pos_embed.permute(0, 2, 3, 1).view(1, -1, dim)
    return torch.cat((class_pos_embed.unsqueeze(0), patch_pos_embed), dim=1)


@torch.no_grad()
def key_facets(model, batch, layers, stride):
    """{layer: key facet [B, heads, T, d]} of one forward (dino.py:172-176, 211-229); the blocks behind the last requested
    layer do not influence it and are not run."""
    p = model.patch_embed.patch_size
    B, _, H, W = batch.shape
    x = nn.functional.conv2d(batch, model.patch_embed.proj.weight, model.patch_embed.proj.bias, stride=stride).flatten(2).transpose(1, 2)
    x = torch.cat((model.cls_token.expand(B, -1, -1), x), dim=1)
    x = x + fix_pos_enc(model, x.shape[1], H, W, p, (stride, stride))
    out = {}
    for i, blk in enumerate(model.blocks):
        if i in layers:
            inp = blk.norm1(x)
            Bn, N, C = inp.shape
            qkv = blk.attn.qkv(inp).reshape(Bn, N, 3, blk.attn.num_heads, C // blk.attn.num_heads).permute(2, 0, 3, 1, 4)
            out[i] = qkv[1]
        if i >= max(layers):
            break
        x = blk(x)
    return out


@torch.no_grad()
def extract_descriptors(model, batch, layers=11, stride=None, resize_shape=None):
    """ViTExtractor.extract_descriptors(batch, layers, facet='key', include_cls=False, resize_shape) (dino.py:279-322)."""
    p = model.patch_embed.patch_size
    stride = stride or p
    B, _, H, W = batch.shape
    num_patches = (1 + (H - p) // stride, 1 + (W - p) // stride)

    def to_map(x):  # [B, heads, T, d] -> [B, d*heads, hp, wp]
        x = x[:, :, 1:, :]
        b = x.shape[0]
        x = x.permute(0, 2, 3, 1).flatten(start_dim=-2, end_dim=-1).unsqueeze(dim=1)
        return x.view(b, 1, num_patches[0], num_patches[1], -1).squeeze(1).permute(0, 3, 1, 2)

    if type(layers) == int:
        x = to_map(key_facets(model, batch, [layers], stride)[layers])
        size = (H // stride, W // stride) if resize_shape is None else resize_shape
        return nn.functional.interpolate(x, size, mode="bilinear")
    feats = key_facets(model, batch, list(layers), stride)
    desc = [to_map(feats[i]) for i in layers]
    if resize_shape is not None:
        desc = [nn.functional.interpolate(x, resize_shape, mode="bilinear") for x in desc]
    return desc
