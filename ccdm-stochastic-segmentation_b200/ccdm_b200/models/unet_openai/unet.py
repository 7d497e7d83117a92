"""UNetModel: the reference's constructor surface and checkpoint layout over the CUDA engine.

Mirrors ``UNetModel`` of /root/reference/ddpm/models/unet_openai/unet.py:402-808:
same constructor arguments (:433-457), same public attributes, same
``state_dict()`` key names and tensor shapes (so ``{"model": ..., "average_model":
...}`` checkpoints load strictly, trainer.py:357-376 / eval_cdm.py:131-144), same
``forward(x, input_condition, feature_condition, timesteps, y=None)`` contract
(:744-808).  It is *not* a stack of ``torch.nn`` layers: the constructor derives an
architecture table (``UNetArch``), registers bare parameter holders under the
reference's module paths, and ``forward`` hands the table to
``ccdm_b200.engine`` which runs hand-written sm_100a kernels.  There is no
PyTorch/CPU fallback: without a B200 and the built library ``forward`` raises.
"""
import math
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import torch
from torch import nn

GN_GROUPS = 32  # nn.py:93-100  normalization(channels) = GroupNorm32(32, channels)


@dataclass
class Layer:
    kind: str                 # 'conv_in' | 'res' | 'attn' | 'down' | 'up'
    path: str                 # module path == state_dict key prefix
    cin: int
    cout: int
    heads: int = 0
    skip_conv: bool = False   # ResBlock.skip_connection is a 1x1 conv (unet.py:227-228)
    emb_off: int = -1         # column of this ResBlock in the fused embedding table


@dataclass
class Block:
    stage: str                # 'in' | 'mid' | 'out'
    index: int
    layers: List[Layer] = field(default_factory=list)
    feat_concat: int = 0      # channels of feature_condition concatenated in front (unet.py:770-788)
    skip_channels: int = 0    # channels popped from the skip stack and concatenated (unet.py:797)


@dataclass
class UNetArch:
    blocks: List[Block]
    params: List[Tuple[str, Tuple[int, ...], str]]  # (key, shape, init)
    emb_cols: int
    head_path: str = "out"


def plan_unet(in_channels, model_channels, out_channels, num_res_blocks, attention_resolutions, channel_mult,
              num_heads, num_head_channels, num_heads_upsample, feature_cond_encoder, feature_condition_idx) -> UNetArch:
    """Architecture table equivalent to the module tree built at unet.py:504-713."""
    params: List[Tuple[str, Tuple[int, ...], str]] = []
    blocks: List[Block] = []
    emb_dim = 4 * model_channels
    emb_cols = 0

    def conv(path, cin, cout, k, init="default", dims=2):
        params.append((path + ".weight", (cout, cin) + (k,) * dims, init))
        params.append((path + ".bias", (cout,), "zeros" if init == "zeros" else "bias:%d" % (cin * k ** dims)))

    def norm(path, c):
        if c % GN_GROUPS:
            raise ValueError(f"GroupNorm(32, {c}): channels must be a multiple of 32 ({path})")
        params.append((path + ".weight", (c,), "ones"))
        params.append((path + ".bias", (c,), "zeros"))

    def n_heads(ch, default_heads):
        if num_head_channels == -1:
            return default_heads
        if ch % num_head_channels:
            raise AssertionError(f"q,k,v channels {ch} is not divisible by num_head_channels {num_head_channels}")
        return ch // num_head_channels

    def res(path, cin, cout) -> Layer:
        nonlocal emb_cols
        norm(path + ".in_layers.0", cin)
        conv(path + ".in_layers.2", cin, cout, 3)
        params.append((path + ".emb_layers.1.weight", (cout, emb_dim), "default"))
        params.append((path + ".emb_layers.1.bias", (cout,), "bias:%d" % emb_dim))
        norm(path + ".out_layers.0", cout)
        conv(path + ".out_layers.3", cout, cout, 3, init="zeros")  # zero_module, unet.py:216-218
        skip = cin != cout
        if skip:
            conv(path + ".skip_connection", cin, cout, 1)
        layer = Layer("res", path, cin, cout, skip_conv=skip, emb_off=emb_cols)
        emb_cols += cout
        return layer

    def attn(path, ch, heads) -> Layer:
        norm(path + ".norm", ch)
        conv(path + ".qkv", ch, 3 * ch, 1, dims=1)
        conv(path + ".proj_out", ch, ch, 1, init="zeros", dims=1)  # zero_module, unet.py:300
        return Layer("attn", path, ch, ch, heads=heads)

    # time_embed (unet.py:506-510)
    params.append(("time_embed.0.weight", (emb_dim, model_channels), "default"))
    params.append(("time_embed.0.bias", (emb_dim,), "bias:%d" % model_channels))
    params.append(("time_embed.2.weight", (emb_dim, emb_dim), "default"))
    params.append(("time_embed.2.bias", (emb_dim,), "bias:%d" % emb_dim))

    ch = input_ch = int(channel_mult[0] * model_channels)
    conv("input_blocks.0.0", in_channels, ch, 3)
    blocks.append(Block("in", 0, [Layer("conv_in", "input_blocks.0.0", in_channels, ch)]))
    chans = [ch]
    ds = 1
    idx = 1
    for level, mult in enumerate(channel_mult):
        for _ in range(num_res_blocks):
            blk = Block("in", idx)
            if feature_cond_encoder is not None and idx in feature_condition_idx and feature_cond_encoder["output_stride"] == ds:
                blk.feat_concat = int(feature_cond_encoder["channels"])  # unet.py:545-550
                ch = ch + blk.feat_concat
            cout = int(mult * model_channels)
            blk.layers.append(res(f"input_blocks.{idx}.0", ch, cout))
            ch = cout
            if ds in attention_resolutions:
                blk.layers.append(attn(f"input_blocks.{idx}.1", ch, n_heads(ch, num_heads)))
            blocks.append(blk)
            chans.append(ch)
            idx += 1
        if level != len(channel_mult) - 1:
            conv(f"input_blocks.{idx}.0.op", ch, ch, 3)  # Downsample, unet.py:136-139
            blocks.append(Block("in", idx, [Layer("down", f"input_blocks.{idx}.0.op", ch, ch)]))
            chans.append(ch)
            idx += 1
            ds *= 2

    mid = Block("mid", 0)
    mid.layers.append(res("middle_block.0", ch, ch))
    mid.layers.append(attn("middle_block.1", ch, n_heads(ch, num_heads)))
    mid.layers.append(res("middle_block.2", ch, ch))
    blocks.append(mid)

    oidx = 0
    for level, mult in list(enumerate(channel_mult))[::-1]:
        for i in range(num_res_blocks + 1):
            ich = chans.pop()
            blk = Block("out", oidx, skip_channels=ich)
            cout = int(model_channels * mult)
            blk.layers.append(res(f"output_blocks.{oidx}.0", ch + ich, cout))
            ch = cout
            j = 1
            if ds in attention_resolutions:
                blk.layers.append(attn(f"output_blocks.{oidx}.{j}", ch, n_heads(ch, num_heads_upsample)))
                j += 1
            if level and i == num_res_blocks:
                conv(f"output_blocks.{oidx}.{j}.conv", ch, ch, 3)  # Upsample, unet.py:103-104
                blk.layers.append(Layer("up", f"output_blocks.{oidx}.{j}.conv", ch, ch))
                ds //= 2
            blocks.append(blk)
            oidx += 1

    norm("out.0", ch)
    conv("out.2", input_ch, out_channels, 3, init="zeros")  # zero_module, unet.py:705
    if ch != input_ch:
        raise ValueError("output head expects the first level's channel count")
    return UNetArch(blocks, params, emb_cols)


class _Holder(nn.Module):
    """A node of the parameter tree: holds ``weight``/``bias`` and/or child nodes."""

    def extra_repr(self):
        return ", ".join(f"{n}{tuple(p.shape)}" for n, p in self._parameters.items())


def _init_param(shape, init):
    t = torch.empty(shape, dtype=torch.float32)
    if init == "ones":
        return t.fill_(1.0)
    if init == "zeros":
        return t.zero_()
    if init.startswith("bias:"):
        bound = 1.0 / math.sqrt(int(init[5:]))
        return t.uniform_(-bound, bound)
    fan_in = 1
    for s in shape[1:]:
        fan_in *= s
    bound = 1.0 / math.sqrt(fan_in)  # torch's default conv/linear init: U(-1/sqrt(fan_in), 1/sqrt(fan_in))
    return t.uniform_(-bound, bound)


class UNetModel(nn.Module):
    """Drop-in for the reference ``UNetModel`` (unet.py:402).  See module docstring."""

    def __init__(self, in_channels, model_channels, out_channels, num_res_blocks, cond_encoded_shape,
                 attention_resolutions, dropout=0, channel_mult=(1, 2, 4, 8), conv_resample=True, dims=2,
                 num_classes=None, use_checkpoint=False, use_fp16=False, num_heads=1, num_head_channels=-1,
                 num_heads_upsample=-1, use_scale_shift_norm=False, resblock_updown=False,
                 use_new_attention_order=False, softmax_output=True, ce_head=False, feature_cond_encoder=None):
        super().__init__()
        # Switches no shipped configuration enables: accepted by
        # the signature, refused loudly when set -- there is no fallback path.
        unsupported = dict(use_fp16=use_fp16, use_scale_shift_norm=use_scale_shift_norm, resblock_updown=resblock_updown,
                           use_new_attention_order=use_new_attention_order, ce_head=ce_head)
        for name, val in unsupported.items():
            if val:
                raise NotImplementedError(f"UNetModel({name}=True) is not implemented by the B200 sampler")
        if num_classes is not None:
            raise NotImplementedError("class-conditional UNet (num_classes) is not implemented by the B200 sampler")
        if dims != 2 or not conv_resample:
            raise NotImplementedError("only dims=2 with learned (conv) resampling is implemented")
        if num_heads_upsample == -1:
            num_heads_upsample = num_heads

        self.in_channels = in_channels
        self.model_channels = model_channels
        self.out_channels = out_channels
        self.num_res_blocks = num_res_blocks
        self.attention_resolutions = attention_resolutions
        self.dropout = dropout
        self.channel_mult = channel_mult
        self.conv_resample = conv_resample
        self.num_classes = num_classes
        self.use_checkpoint = use_checkpoint
        self.dtype = torch.float32
        self.num_heads = num_heads
        self.num_head_channels = num_head_channels
        self.num_heads_upsample = num_heads_upsample
        self.cond_encoded_shape = cond_encoded_shape
        self.sofmtax_output = softmax_output  # (sic) attribute name of the reference, unet.py:478
        self.use_ce_head = ce_head
        self.out_ce = None
        self.feature_cond_encoder = feature_cond_encoder
        if feature_cond_encoder is not None:
            if feature_cond_encoder["type"] != "dino":
                raise NotImplementedError(f"{feature_cond_encoder['type']} not implemented")
            if feature_cond_encoder["scale"] != "single":
                raise NotImplementedError(f"feature_cond_encoder {feature_cond_encoder['type']} with scale"
                                          f" {feature_cond_encoder['scale']} not implemented")
            tl = feature_cond_encoder["target_layer"]
            self.feature_condition_idx = [tl] if tl is not None else [None]
        else:
            self.feature_condition_idx = []

        self.arch = plan_unet(in_channels, model_channels, out_channels, num_res_blocks, attention_resolutions,
                              channel_mult, num_heads, num_head_channels, num_heads_upsample, feature_cond_encoder,
                              self.feature_condition_idx)
        for key, shape, init in self.arch.params:
            *path, leaf = key.split(".")
            node = self
            for name in path:
                if name not in node._modules:
                    node.add_module(name, _Holder())
                node = node._modules[name]
            node.register_parameter(leaf, nn.Parameter(_init_param(shape, init)))
        self._engines = {}

    # -- reference API -------------------------------------------------------------------
    def convert_to_fp16(self):
        raise NotImplementedError("use DenoisingModel.precision = 'bf16' for reduced-precision storage")

    def convert_to_fp32(self):
        return None

    @property
    def feat_channels(self) -> int:
        return sum(b.feat_concat for b in self.arch.blocks)

    def engine(self, precision: str = "exact", dry_run: bool = False):
        from ...engine import UNetEngine
        key = (precision, dry_run, next(self.parameters()).device)
        eng = self._engines.get(key)
        if eng is None:
            eng = self._engines[key] = UNetEngine(self, precision, dry_run=dry_run)
        return eng

    @torch.no_grad()
    def forward(self, x, input_condition, feature_condition, timesteps, y=None):
        """One denoiser evaluation (unet.py:744-808): returns ``{"diffusion_out": softmax [B,K,H,W], "logits": None}``.

        ``x`` must be a one-hot label map (what every caller on the hot path passes:
        x_t of the chain, q(x_t|x_0) samples in training/validation); it is reduced to
        uint8 labels and the input concat+conv runs on those.  Inference only.
        """
        assert (y is not None) == (self.num_classes is not None), \
            "must specify y if and only if the model is class-conditional"
        precision = getattr(self, "precision", "exact")
        probs = self.engine(precision).single_step(x, input_condition, feature_condition, timesteps,
                                                   softmax=self.sofmtax_output)
        return {"diffusion_out": probs, "logits": None}
