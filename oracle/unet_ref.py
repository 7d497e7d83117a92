"""Plain-PyTorch fp32 restatement of the reference UNet forward, driven by a state_dict.

TEST INFRASTRUCTURE (see package docstring).  The network structure is inferred
from the reference's checkpoint key names alone (SURVEY.md 8b "Weights"), so this
file shares no construction code with the product: every ``input_blocks.N.M.*``
group is classified by the parameter names it holds.

Restates /root/reference/ddpm/models/unet_openai/unet.py:744-808 (forward),
:242-262 (ResBlock), :305-311 + :343-360 (AttentionBlock / QKVAttentionLegacy),
:106-116 (Upsample), :144-146 (Downsample), nn.py:103-121 (timestep_embedding),
nn.py:17-19,93-100 (GroupNorm32 with 32 groups).
"""
import math
import re

import torch
import torch.nn.functional as F


def _timestep_embedding(timesteps, dim, max_period=10000):
    # nn.py:103-121
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half)
    args = timesteps[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def _gn(sd, prefix, x):
    return F.group_norm(x.float(), 32, sd[prefix + ".weight"], sd[prefix + ".bias"], eps=1e-5)


def _resblock(sd, p, x, emb):
    # unet.py:242-262 (no up/down, no scale-shift: the shipped configuration)
    h = F.conv2d(F.silu(_gn(sd, p + ".in_layers.0", x)), sd[p + ".in_layers.2.weight"], sd[p + ".in_layers.2.bias"], padding=1)
    e = F.linear(F.silu(emb), sd[p + ".emb_layers.1.weight"], sd[p + ".emb_layers.1.bias"])
    h = h + e[:, :, None, None]
    h = F.conv2d(F.silu(_gn(sd, p + ".out_layers.0", h)), sd[p + ".out_layers.3.weight"], sd[p + ".out_layers.3.bias"], padding=1)
    if p + ".skip_connection.weight" in sd:
        w = sd[p + ".skip_connection.weight"]
        x = F.conv2d(x, w, sd[p + ".skip_connection.bias"], padding=w.shape[-1] // 2)
    return x + h


def _attention(sd, p, x, head_channels, num_heads):
    # unet.py:305-311, 343-360 (legacy order: heads split first, then q|k|v)
    b, c, hh, ww = x.shape
    xf = x.reshape(b, c, -1)
    qkv = F.conv1d(_gn(sd, p + ".norm", xf), sd[p + ".qkv.weight"], sd[p + ".qkv.bias"])
    n_heads = c // head_channels if head_channels > 0 else num_heads
    ch = c // n_heads
    q, k, v = qkv.reshape(b * n_heads, ch * 3, -1).split(ch, dim=1)
    scale = 1 / math.sqrt(math.sqrt(ch))
    w = torch.einsum("bct,bcs->bts", q * scale, k * scale)
    w = torch.softmax(w.float(), dim=-1)
    a = torch.einsum("bts,bcs->bct", w, v).reshape(b, -1, hh * ww)
    a = F.conv1d(a, sd[p + ".proj_out.weight"], sd[p + ".proj_out.bias"])
    return (xf + a).reshape(b, c, hh, ww)


def _run_group(sd, prefix, h, emb, head_channels, num_heads, taps=None):
    j = 0
    while True:
        p = f"{prefix}.{j}"
        if p + ".in_layers.0.weight" in sd:
            h = _resblock(sd, p, h, emb)
        elif p + ".qkv.weight" in sd:
            h = _attention(sd, p, h, head_channels, num_heads)
        elif p + ".op.weight" in sd:
            h = F.conv2d(h, sd[p + ".op.weight"], sd[p + ".op.bias"], stride=2, padding=1)
        elif p + ".conv.weight" in sd:
            h = F.interpolate(h, scale_factor=2, mode="nearest")
            h = F.conv2d(h, sd[p + ".conv.weight"], sd[p + ".conv.bias"], padding=1)
        elif p + ".weight" in sd:
            h = F.conv2d(h, sd[p + ".weight"], sd[p + ".bias"], padding=1)
        else:
            break
        if taps is not None:
            taps[p] = h
        j += 1
    return h


def _count(sd, stem):
    idx = {int(m.group(1)) for k in sd for m in [re.match(rf"{stem}\.(\d+)\.", k)] if m}
    return max(idx) + 1 if idx else 0


@torch.no_grad()
def unet_forward(sd, x, image, feature_condition, timesteps, head_channels=32, num_heads=1,
                 feature_condition_idx=None, softmax_output=True, taps=None):
    """``sd``: UNet state_dict with the reference's key names (fp32 CPU tensors).

    x: one-hot [B,K,H,W]; image: [B,C_img,H,W]; feature_condition: None or
    [B,C_f,H/8,W/8] concatenated in front of ``input_blocks[feature_condition_idx]``
    (unet.py:770-788); timesteps: [B] float.  Returns softmax probabilities
    [B,K,H,W] (unet.py:701-707,803).  ``taps`` (dict) collects every layer output.
    """
    sd = {k: v.float() for k, v in sd.items()}
    model_channels = sd["time_embed.0.weight"].shape[1]
    emb = _timestep_embedding(timesteps, model_channels)
    emb = F.linear(emb, sd["time_embed.0.weight"], sd["time_embed.0.bias"])
    emb = F.linear(F.silu(emb), sd["time_embed.2.weight"], sd["time_embed.2.bias"])
    h = torch.cat([x.float(), image.float()], dim=1)
    hs = []
    for i in range(_count(sd, "input_blocks")):
        if feature_condition is not None and feature_condition_idx is not None and i == feature_condition_idx:
            h = torch.cat([h, feature_condition.float()], dim=1)
        h = _run_group(sd, f"input_blocks.{i}", h, emb, head_channels, num_heads, taps)
        hs.append(h)
    h = _run_group(sd, "middle_block", h, emb, head_channels, num_heads, taps)
    for i in range(_count(sd, "output_blocks")):
        h = torch.cat([h, hs.pop()], dim=1)
        h = _run_group(sd, f"output_blocks.{i}", h, emb, head_channels, num_heads, taps)
    logits = F.conv2d(F.silu(_gn(sd, "out.0", h)), sd["out.2.weight"], sd["out.2.bias"], padding=1)
    if taps is not None:
        taps["logits"] = logits
    return torch.softmax(logits, dim=1) if softmax_output else logits
