"""Host side of the DINO ViT condition encoder: weight packing and the launch sequence.

The reference runs ``ViTExtractor.extract_descriptors`` (ddpm/models/dino.py:279-309) once per image, outside the T loop:
one forward of a torch.hub ViT with a hook that recomputes ``qkv(norm1(x))`` of block ``layer`` and keeps the key part.
Here the blocks up to ``layer`` run as hand-written sm_100a kernels (``csrc/vit.cu`` + the sampler's own attention kernel) in
the fp16x2 storage format, and the blocks behind ``layer`` -- which cannot influence the descriptor -- are not run at all.

torch is used for device memory and the current stream only.
"""
import ctypes
import math
from typing import Dict, List, Optional, Sequence, Union

import torch

from . import _lib
from ._lib import Op
from .engine import pack_conv_weight_x3

LN_EPS = 1e-6  # vision_transformer.py: norm_layer = partial(nn.LayerNorm, eps=1e-6)


class ViTEngine:
    def __init__(self, vit, stride: int):
        _lib.require_device()
        p0 = next(vit.parameters())
        if p0.device.type != "cuda":
            raise _lib.CcdmError("the DINO encoder must live on a CUDA device (B200); there is no CPU path")
        self.vit = vit
        self.device = p0.device
        self.patch = int(vit.patch_embed.patch_size)
        self.stride = int(stride)
        if self.patch % self.stride:
            raise AssertionError(f"stride {self.stride} should divide patch_size {self.patch}")  # dino.py:131-132
        self.D = int(vit.embed_dim)
        self.heads = int(vit.num_heads)
        self.hd = self.D // self.heads
        if self.hd != 64 or self.D not in (384, 768):
            raise NotImplementedError(f"DINO encoder: embed_dim {self.D} / heads {self.heads} (ViT-S and ViT-B: head size 64)")
        self.depth = len(vit.blocks)
        self.w: Dict[str, torch.Tensor] = {}
        self.shift = 0
        self._stamp = None
        self._pos: Dict[tuple, torch.Tensor] = {}
        self._ws: Dict[tuple, Dict[str, torch.Tensor]] = {}

    # -- weights ------------------------------------------------------------------------------------
    def _current_stamp(self):
        return tuple((p.data_ptr(), p._version) for p in self.vit.parameters())

    def invalidate(self):
        self._stamp = None

    def refresh(self) -> bool:
        """(Re)pack when a parameter changed (load_state_dict, .to(), in-place writes): same rule as engine.PackedWeights."""
        stamp = self._current_stamp()
        if self.w and stamp == self._stamp:
            return False
        L = _lib.lib()
        nt = int(L.ccdm_vit_linear_nt())
        sd = {k: v.detach().to(device=self.device, dtype=torch.float32) for k, v in self.vit.state_dict().items()}
        D, heads, hd = self.D, self.heads, self.hd
        wmax = max([1e-30] + [float(v.abs().max()) for k, v in sd.items() if k.startswith("blocks.") and v.dim() == 2])
        self.shift = int(max(0, min(13, math.floor(math.log2(32768.0 / wmax)))))
        w: Dict[str, torch.Tensor] = {}

        def lin(name, weight, bias):
            w[name + ".w"] = pack_conv_weight_x3(weight.reshape(weight.shape[0], weight.shape[1], 1, 1), self.shift, nt).reshape(-1).contiguous()
            w[name + ".b"] = bias.contiguous()

        # qkv rows [which][head][d] (Attention.forward: reshape(B, N, 3, heads, d)) -> the attention kernel's per-head
        # q|k|v order [head][which][d] (QKVAttentionLegacy layout)
        perm = torch.arange(3 * D, device=self.device).reshape(3, heads, hd).permute(1, 0, 2).reshape(-1)
        w["patch.wt"] = sd["patch_embed.proj.weight"].reshape(D, -1).t().contiguous()
        w["patch.b"] = sd["patch_embed.proj.bias"].contiguous()
        w["cls"] = sd["cls_token"].reshape(D).contiguous()
        w["pos"] = sd["pos_embed"].reshape(-1, D).contiguous()
        for i in range(self.depth):
            p = "blocks.%d." % i
            for n in ("norm1", "norm2"):
                w[p + n + ".g"] = sd[p + n + ".weight"].contiguous()
                w[p + n + ".b"] = sd[p + n + ".bias"].contiguous()
            qw, qb = sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"]
            lin(p + "qkv", qw[perm], qb[perm])
            lin(p + "key", qw[D:2 * D], qb[D:2 * D])  # the hook's facet (dino.py:172-176): rows D .. 2D, channel = h * d + i
            lin(p + "proj", sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"])
            lin(p + "fc1", sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])
            lin(p + "fc2", sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])
        self.w = w
        self._pos.clear()
        self._stamp = stamp
        return True

    def _sp(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def pos_embed(self, H: int, W: int, hp: int, wp: int) -> torch.Tensor:
        """interpolate_pos_encoding (dino.py:86-117 / the hub model's own method): [1 + hp*wp, D]."""
        key = (H, W, hp, wp)
        t = self._pos.get(key)
        if t is not None:
            return t
        pos = self.w["pos"]
        N = pos.shape[0] - 1
        if hp * wp == N and H == W:
            t = pos
        else:
            n = int(math.sqrt(N))
            if n * n != N:
                raise ValueError(f"pos_embed has {N} patch positions: not a square grid")
            t = torch.empty(1 + hp * wp, self.D, dtype=torch.float32, device=self.device)
            _lib.check(_lib.lib().ccdm_vit_pos_embed(pos.data_ptr(), n, self.D, hp, wp, (hp + 0.1) / math.sqrt(N), (wp + 0.1) / math.sqrt(N),
                                                     t.data_ptr(), self._sp()), "vit_pos_embed")
        self._pos[key] = t
        return t

    def _workspace(self, B: int, T: int) -> Dict[str, torch.Tensor]:
        key = (B, T)
        ws = self._ws.get(key)
        if ws is None:
            if len(self._ws) >= 2:
                self._ws.clear()
            D = self.D

            def buf(c):
                return torch.empty(B * c * T * 2, dtype=torch.float16, device=self.device)

            ws = self._ws[key] = dict(x0=buf(D), x1=buf(D), n=buf(D), qkv=buf(3 * D), a=buf(D), h=buf(4 * D), k=buf(D))
        return ws

    # -- the launch sequence ------------------------------------------------------------------------
    @torch.no_grad()
    def key_descriptors(self, batch: torch.Tensor, layers: Sequence[int], out_sizes: Sequence[Optional[tuple]],
                        tokens_out: Optional[dict] = None, timeline: Optional[list] = None) -> List[torch.Tensor]:
        """One pass over blocks 0 .. max(layers); returns, per requested layer, its key facet as a descriptor map
        [B, D, Ho, Wo] fp32 NCHW (``out_sizes[i]`` or the patch grid).  ``tokens_out``: optional dict that receives raw copies
        of intermediate token tensors (tests); ``timeline``: optional list that receives (kernel kind, start event, end event)
        per launch (bench.py's per-kernel shares)."""
        L = _lib.lib()
        if batch.dim() != 4 or batch.shape[1] != 3:
            raise ValueError(f"the encoder takes [B, 3, H, W] images; got {tuple(batch.shape)}")
        if not layers or min(layers) < 0 or max(layers) >= self.depth:
            raise ValueError(f"layers {list(layers)}: a number between 0 and {self.depth - 1}")
        B, _, H, W = batch.shape
        p, s, D = self.patch, self.stride, self.D
        if H < p or W < p:
            raise ValueError(f"image {H}x{W} is smaller than one patch ({p})")
        hp, wp = 1 + (H - p) // s, 1 + (W - p) // s
        T = 1 + hp * wp
        self.refresh()
        w, sp = self.w, self._sp()
        img = batch.to(device=self.device, dtype=torch.float32).contiguous()
        ws = self._workspace(B, T)
        pos = self.pos_embed(H, W, hp, wp)
        acc_shift = self.shift + _lib.F16X2_SCALE_LOG2

        def launch(kind, rc_fn):
            """One kernel launch; with a ``timeline`` it is bracketed by two events."""
            if timeline is not None:
                ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ea.record()
            _lib.check(rc_fn(), kind)
            if timeline is not None:
                eb.record()
                timeline.append((kind, ea, eb))

        def linear(x, name, out, cin, cout, gelu=0, res=None):
            launch("vit_linear %d->%d" % (cin, cout), lambda: L.ccdm_vit_linear(
                x.data_ptr(), w[name + ".w"].data_ptr(), w[name + ".b"].data_ptr(), res.data_ptr() if res is not None else None, B, T,
                cin, cout, gelu, acc_shift, out.data_ptr(), sp))

        def layernorm(x, name, out):
            launch("vit_layernorm", lambda: L.ccdm_vit_layernorm(x.data_ptr(), w[name + ".g"].data_ptr(), w[name + ".b"].data_ptr(), B, T, D,
                                                                 LN_EPS, out.data_ptr(), sp))

        launch("vit_patch_embed", lambda: L.ccdm_vit_patch_embed(img.data_ptr(), w["patch.wt"].data_ptr(), w["patch.b"].data_ptr(),
                                                                 w["cls"].data_ptr(), pos.data_ptr(), B, H, W, p, s, D, ws["x0"].data_ptr(), sp))
        att = Op(kind=_lib.OP_ATTENTION, dtype=_lib.DT_F16X2, out_dtype=_lib.DT_F16X2, B=B, Hin=1, Win=T, Hout=1, Wout=T, C0=3 * D,
                 Cout=D, heads=self.heads, head_dim=self.hd, exact=0, src0=ws["qkv"].data_ptr(), out=ws["a"].data_ptr())
        results: Dict[int, torch.Tensor] = {}
        x, y = ws["x0"], ws["x1"]
        if tokens_out is not None:
            tokens_out["tokens"] = x.clone()
        last = max(layers)
        for i in range(last + 1):
            pre = "blocks.%d." % i
            layernorm(x, pre + "norm1", ws["n"])
            if i in layers:
                linear(ws["n"], pre + "key", ws["k"], D, D)
                size = out_sizes[list(layers).index(i)] or (hp, wp)
                out = torch.empty(B, D, size[0], size[1], dtype=torch.float32, device=self.device)
                launch("vit_descriptor", lambda: L.ccdm_vit_descriptor(ws["k"].data_ptr(), B, T, self.heads, self.hd, hp, wp, size[0], size[1],
                                                                       out.data_ptr(), sp))
                results[i] = out
            if i == last:
                break
            linear(ws["n"], pre + "qkv", ws["qkv"], D, 3 * D)
            launch("attention T=%d heads=%d d=%d" % (T, self.heads, self.hd), lambda: L.ccdm_launch_op(ctypes.byref(att), sp))
            linear(ws["a"], pre + "proj", y, D, D, res=x)            # x + proj(attn(norm1(x)))
            layernorm(y, pre + "norm2", ws["n"])
            linear(ws["n"], pre + "fc1", ws["h"], D, 4 * D, gelu=1)
            linear(ws["h"], pre + "fc2", x, 4 * D, D, res=y)         # ... + fc2(gelu(fc1(norm2(.))))
            if tokens_out is not None:
                tokens_out["block%d" % i] = x.clone()
        img.record_stream(torch.cuda.current_stream(self.device))
        return [results[i] for i in layers]


def tokens_to_float(raw: torch.Tensor, B: int, C: int, T: int) -> torch.Tensor:
    """fp16x2 token tensor [B][C/8][2][T][8] -> fp32 [B, T, C] (tests)."""
    v = raw.view(B, C // 8, 2, T, 8).float()
    return ((v[:, :, 0] + v[:, :, 1]) * (1.0 / float(2 ** _lib.F16X2_SCALE_LOG2))).permute(0, 2, 1, 3).reshape(B, T, C).contiguous()


def tokens_from_float(x: torch.Tensor) -> torch.Tensor:
    """fp32 [B, T, C] -> fp16x2 token tensor [B][C/8][2][T][8] (tests)."""
    from .engine import split_f16x2
    B, T, C = x.shape
    hi, lo = split_f16x2(x)
    both = torch.stack([hi, lo], 0).reshape(2, B, T, C // 8, 8)  # [part, B, T, G, 8]
    return both.permute(1, 3, 0, 2, 4).contiguous()
