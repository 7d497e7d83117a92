"""Host side of the CUDA engine: compiles a ``UNetArch`` into a launch program and runs chains.

What lives here (all host logic; every FLOP is in ``libccdm_b200.so``):

* ``pack_weights``  -- repack the reference-layout parameters (OIHW fp32, checkpoint key
  names) into kernel layouts (``[tap][Cin][Cout]``, fused biases, the concatenated
  timestep-embedding projection) inside ONE device buffer whose addresses never change,
  re-done lazily whenever a parameter is modified (``load_state_dict`` / ``.to()`` /
  in-place EMA writes, SURVEY.md 8b "staleness hazard").
* ``Program``       -- the op list of one reverse step for a fixed (B, H, W): every
  activation gets an offset in a liveness-planned workspace; the op structs carry
  absolute device addresses so the C side can capture them into a CUDA graph once.
* ``UNetEngine``    -- caches programs, runs ``single_step`` (reference
  ``UNetModel.forward``, unet.py:744-808) and ``run_chain`` (reference
  ``DenoisingModel.forward_denoising``, diffusion_denoising.py:164-215).

torch is used for device memory, streams and the global generator (noise parity
with the reference) only.
"""
import ctypes
import logging
import math
import os
from collections import OrderedDict
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import torch

from . import _lib
from ._lib import Op, StepEntry

ALIGN = 256
LOGGER = logging.getLogger(__name__)

# precision modes of the engine -> (activation storage CCDM_DT_*, bytes per stored element, tensor-core kernels?)
#   fp32  : NHWC fp32, FFMA kernels                      (the in-library reference every other mode is tested against)
#   exact : fp16x2 (hi + lo planes), tcgen05 3-MMA split  (fp32-grade products at tensor-core speed; the default)
#   bf16  : bf16 planes, tcgen05                           (fast mode)
PRECISIONS = {"fp32": (_lib.DT_F32, 4, False), "exact": (_lib.DT_F16X2, 4, True), "bf16": (_lib.DT_BF16, 2, True)}
PROGRAM_CACHE = 4  # programs (workspaces of ~1 GB at the benchmark shapes) kept per engine, least recently used first out


def _ceil(a, b):
    return (a + b - 1) // b * b


# --------------------------------------------------------------------------------------
# workspace planning
# --------------------------------------------------------------------------------------
@dataclass
class Ten:
    name: str
    C: int
    H: int
    W: int
    esize: int            # bytes per element
    want_stat: bool
    fmt: int = 0          # CCDM_DT_* storage format
    nbytes: int = 0
    off: int = -1         # byte offset in the activation arena
    stat_off: int = -1    # byte offset in the stat arena
    born: int = -1
    last: int = -1
    external: bool = False  # lives outside the arena (feature condition)
    addr: int = 0
    stat_addr: int = 0
    part_addr: int = 0            # deferred-fold statistics: the producer's per-CTA partial rows (bf16 / tensor-core mode)
    stat_layout: Optional[tuple] = None  # (slots, items per sample, items, grid, row length) of those rows


class Arena:
    """First-fit allocator over op-index lifetimes (tensors die after their last reader)."""

    def __init__(self):
        self.live = []  # (off, nbytes, ten)
        self.size = 0

    def alloc(self, t: Ten):
        self.live.sort(key=lambda x: x[0])
        pos = 0
        for off, nb, _ in self.live:
            if off - pos >= t.nbytes:
                break
            pos = max(pos, off + nb)
        t.off = pos
        self.live.append((pos, t.nbytes, t))
        self.size = max(self.size, pos + t.nbytes)

    def release_dead(self, op_index: int):
        self.live = [x for x in self.live if x[2].last > op_index]


# --------------------------------------------------------------------------------------
# weights
# --------------------------------------------------------------------------------------
def pack_conv_weight(w: torch.Tensor, cin_pad: Optional[int] = None) -> torch.Tensor:
    """OIHW (or OI1 for Conv1d) -> the kernels' [tap][CinP][CoutP] fp32 layout, zero padded
    (CinP defaults to Cin, CoutP = ceil32(Cout))."""
    if w.dim() == 3:
        w = w[:, :, :, None]
    co, ci, kh, kw = w.shape
    cip = cin_pad or ci
    out = torch.zeros(kh * kw, cip, _ceil(co, 32), dtype=torch.float32, device=w.device)
    out[:, :ci, :co] = w.float().permute(2, 3, 1, 0).reshape(kh * kw, ci, co)
    return out


# A/B switch: identity residuals of the bf16 mode as an extra 1x1 MMA chunk with identity weights (default) or as a
# read-and-add in the epilogue (CCDM_IDENT_SKIP=0)
_IDENT_SKIP = os.environ.get("CCDM_IDENT_SKIP", "1") != "0"


def pack_conv_weight_tc(w: torch.Tensor, nt: Optional[int] = None) -> torch.Tensor:
    """OIHW (or OI1) -> the tcgen05 kernel's bf16 [cc][Cin/8][tap][NT][8] layout (cc = chunk of NT output
    channels of ceil16(Cout)): every (chunk, 8-channel plane, tap) is NT rows of 16 bytes = the UMMA K-major
    no-swizzle canonical form, and any run of planes of one chunk is contiguous (one bulk copy per K chunk)."""
    if w.dim() == 3:
        w = w[:, :, :, None]
    co, ci, kh, kw = w.shape
    assert ci % 8 == 0
    if nt is None:
        nt = int(_lib.lib().ccdm_conv_tc_nt(co, kh * kw, 0))
    cop = _ceil(co, 16)
    assert cop % nt == 0
    taps = kh * kw
    full = torch.zeros(cop, ci, taps, dtype=torch.float32, device=w.device)
    full[:co] = w.float().reshape(co, ci, taps)
    out = full.reshape(cop // nt, nt, ci // 8, 8, taps).permute(0, 2, 4, 1, 3).contiguous()
    return out.to(torch.bfloat16)


def split_f16x2(x: torch.Tensor, scale_log2: int = _lib.F16X2_SCALE_LOG2):
    """fp32 -> (hi, lo) fp16 with hi + lo == 2^scale_log2 * x to 2^-23 relative (CCDM_DT_F16X2), saturating like the kernels."""
    y = (x.float() * float(2 ** scale_log2)).clamp(-65504.0, 65504.0)
    hi = y.to(torch.float16)
    lo = (y - hi.float()).to(torch.float16)
    return hi, lo


def pack_conv_weight_x3(w: torch.Tensor, shift: int, nt: Optional[int] = None) -> torch.Tensor:
    """OIHW (or OI1) -> the fp16x2 tensor-core layout [cc][Cin/8][tap][2][NT][8] fp16 of 2^shift * w: per (chunk, 8-channel
    plane, tap) the NT hi rows, then the NT lo rows."""
    if w.dim() == 3:
        w = w[:, :, :, None]
    co, ci, kh, kw = w.shape
    assert ci % 8 == 0
    if nt is None:
        nt = int(_lib.lib().ccdm_conv_tc_nt(co, kh * kw, 1))
    cop = _ceil(co, 16)
    assert cop % nt == 0
    taps = kh * kw
    full = torch.zeros(cop, ci, taps, dtype=torch.float32, device=w.device)
    full[:co] = w.float().reshape(co, ci, taps)
    hi, lo = split_f16x2(full, shift)
    both = torch.stack([hi, lo], 0)  # [2, cop, ci, taps]
    return both.reshape(2, cop // nt, nt, ci // 8, 8, taps).permute(1, 3, 5, 0, 2, 4).contiguous()


def to_pm_x3(x_nhwc: torch.Tensor) -> torch.Tensor:
    """NHWC fp32 [B,H,W,C] -> fp16x2 plane-major [B, C/8, 2, H, W, 8] (contiguous)."""
    B, H, W, C = x_nhwc.shape
    assert C % 8 == 0
    hi, lo = split_f16x2(x_nhwc)
    both = torch.stack([hi, lo], -1).reshape(B, H, W, C // 8, 8, 2)
    return both.permute(0, 3, 5, 1, 2, 4).contiguous()


def from_pm_x3(x: torch.Tensor) -> torch.Tensor:
    """fp16x2 plane-major [B, C/8, 2, H, W, 8] -> NHWC fp32 [B,H,W,C] = (hi + lo) / 16."""
    B, G, _, H, W, _ = x.shape
    v = (x[:, :, 0].float() + x[:, :, 1].float()) * (1.0 / float(2 ** _lib.F16X2_SCALE_LOG2))
    return v.permute(0, 2, 3, 1, 4).reshape(B, H, W, G * 8).contiguous()


def to_pm(x_nhwc: torch.Tensor) -> torch.Tensor:
    """NHWC [B,H,W,C] -> the bf16 kernels' plane-major layout [B, C/8, H, W, 8] (contiguous)."""
    B, H, W, C = x_nhwc.shape
    assert C % 8 == 0
    return x_nhwc.reshape(B, H, W, C // 8, 8).permute(0, 3, 1, 2, 4).contiguous()


def from_pm(x_pm: torch.Tensor) -> torch.Tensor:
    """plane-major [B, C/8, H, W, 8] -> NHWC [B,H,W,C] (contiguous)."""
    B, G, H, W, _ = x_pm.shape
    return x_pm.permute(0, 2, 3, 1, 4).reshape(B, H, W, G * 8).contiguous()


def subpixel_weights(w: torch.Tensor) -> torch.Tensor:
    """3x3 weights [Cout,Cin,3,3] of ``conv3x3(nearest_x2(x))`` -> [Cout,Cin,4,4]: for every output parity
    p = 2*py+px the 2x2 conv on the LOW-resolution input that produces output pixels (2y+py, 2x+px); tap
    t = 2*ry+rx reads low-res pixel (y-1+py+ry, x-1+px+rx) and carries the sum of the 3x3 weights whose upsampled
    source pixel is that low-res pixel (Upsample, unet.py:106-116; zero padding of the upsampled image coincides
    with zero padding of the low-res image)."""
    rows = {(0, 0): (0,), (0, 1): (1, 2), (1, 0): (0, 1), (1, 1): (2,)}  # (parity, r) -> original taps
    co, ci = w.shape[:2]
    out = torch.zeros(co, ci, 4, 4, dtype=torch.float32, device=w.device)
    wf = w.float()
    for py in (0, 1):
        for px in (0, 1):
            for ry in (0, 1):
                for rx in (0, 1):
                    acc = 0
                    for dy in rows[(py, ry)]:
                        for dx in rows[(px, rx)]:
                            acc = acc + wf[:, :, dy, dx]
                    out[:, :, 2 * py + px, 2 * ry + rx] = acc
    return out


# the DINO channels of input_blocks[10] (unet.py:770-788, 545-550) are the same in every reverse step.  Their
# GroupNorm groups, SiLU and share of the 3x3 conv -- and their share of the 1x1 skip conv -- are computed ONCE per chain into
# two maps that the per-step convs pick up as identity-weight K chunks.  CCDM_FOLD_FEAT=0 keeps the per-step concatenation (A/B).
_FOLD_FEAT = os.environ.get("CCDM_FOLD_FEAT", "1") != "0"


def feat_fold_split(c_h: int, f: int):
    """GroupNorm(32) over cat[h (c_h channels), features (f channels)]: (channels per group, number of leading feature channels
    that share a group with h, that number rounded up to a K chunk of 16)."""
    cpg = (c_h + f) // 32
    n_f = (-c_h) % cpg
    return cpg, n_f, _ceil(n_f, 16)


def pack_bias(b: torch.Tensor) -> torch.Tensor:
    out = torch.zeros(_ceil(b.numel(), 32), dtype=torch.float32, device=b.device)
    out[:b.numel()] = b.float()
    return out


class PackedWeights:
    """All kernel-layout parameters in one flat fp32 device buffer with stable addresses."""

    def __init__(self, unet, device, with_tc: bool = False, x3: bool = False):
        self.unet = unet
        self.device = device
        self.slots: Dict[str, tuple] = {}  # name -> (offset_floats, numel)
        self.total = 0
        self.buf: Optional[torch.Tensor] = None
        self.with_tc = with_tc            # also keep tensor-core layouts of every conv (bf16, or fp16 hi + lo rows when x3)
        self.x3 = x3
        self.shift = 0                    # x3: the packed weights are 2^shift * w (one power of two for the whole model)
        self.slots16: Dict[str, tuple] = {}
        self.total16 = 0
        self.buf16: Optional[torch.Tensor] = None
        self._stamp = None
        self._layout()

    def _reserve(self, name, numel, tc_shape=None):
        self.slots[name] = (self.total, numel)
        self.total += _ceil(numel, ALIGN // 4)
        if tc_shape is not None and self.with_tc:
            taps, ci, co = tc_shape
            n16 = taps * ci * _ceil(co, 16) * (2 if self.x3 else 1)
            self.slots16[name] = (self.total16, n16)
            self.total16 += _ceil(n16, ALIGN // 2)

    def _layout(self):
        arch = self.unet.arch
        mc = self.unet.model_channels
        ed = 4 * mc
        for k, n in (("te0_w", ed * mc), ("te0_b", ed), ("te2_w", ed * ed), ("te2_b", ed)):
            self._reserve(k, n)
        self._reserve("emb_w", arch.emb_cols * ed)
        self._reserve("emb_b", arch.emb_cols)
        for blk in arch.blocks:
            for L in blk.layers:
                p = L.path
                if L.kind == "conv_in":
                    # fp32: [tap][ceil8(Cin)][CoutP]; bf16: a tensor-core conv over ceil16(Cin) zero-padded channels
                    self._reserve(p + ":w", 9 * _ceil(L.cin, 8) * _ceil(L.cout, 32), (9, _ceil(L.cin, 16), L.cout))
                    self._reserve(p + ":b", _ceil(L.cout, 32))
                elif L.kind in ("down", "up"):
                    self._reserve(p + ":w", 9 * L.cin * _ceil(L.cout, 32), (16 if L.kind == "up" else 9, L.cin, L.cout))
                    self._reserve(p + ":b", _ceil(L.cout, 32))
                elif L.kind == "res":
                    self._reserve(p + ":g1", L.cin); self._reserve(p + ":be1", L.cin)
                    self._reserve(p + ":w1", 9 * L.cin * L.cout, (9, L.cin, L.cout)); self._reserve(p + ":b1", L.cout)
                    self._reserve(p + ":g2", L.cout); self._reserve(p + ":be2", L.cout)
                    self._reserve(p + ":w2", 9 * L.cout * L.cout, (9, L.cout, L.cout)); self._reserve(p + ":b2", L.cout)
                    if L.skip_conv:
                        self._reserve(p + ":ws", L.cin * L.cout, (1, L.cin, L.cout))
                    if blk.feat_concat and L is blk.layers[0] and L.skip_conv:
                        # folded form of the feature-concat ResBlock (feat_fold_split): per-step conv over h + the f16 leading
                        # feature channels, per-chain convs over the feature tensor, skip conv over h + the per-chain skip map
                        f = blk.feat_concat
                        c_h = L.cin - f
                        _, _, f16 = feat_fold_split(c_h, f)
                        cs = c_h + f16
                        self._reserve(p + ":g1s", cs); self._reserve(p + ":be1s", cs)
                        self._reserve(p + ":w1s", 9 * cs * L.cout, (9, cs, L.cout))
                        self._reserve(p + ":g1m", f); self._reserve(p + ":be1m", f)
                        self._reserve(p + ":w1m", 9 * f * L.cout, (9, f, L.cout))
                        self._reserve(p + ":wsm", f * L.cout, (1, f, L.cout))
                        self._reserve(p + ":ws2", (c_h + L.cout) * L.cout, (1, c_h + L.cout, L.cout))
                        self._reserve(p + ":b0", _ceil(L.cout, 32))
                elif L.kind == "attn":
                    self._reserve(p + ":g", L.cin); self._reserve(p + ":be", L.cin)
                    self._reserve(p + ":wqkv", L.cin * 3 * L.cin, (1, L.cin, 3 * L.cin)); self._reserve(p + ":bqkv", 3 * L.cin)
                    self._reserve(p + ":wproj", L.cin * L.cin, (1, L.cin, L.cin)); self._reserve(p + ":bproj", L.cin)
        # bf16 mode folds identity residuals (ResBlock skip = Identity, attention x + proj) into the MMA as a 1x1
        # "skip conv" with identity weights: x * I accumulates exactly in fp32 and the epilogue loses a load
        # (the packed layout depends on the N tile of the conv the chunk rides on: a ResBlock's 3x3 conv and an attention
        # block's 1x1 projection of the same width use different tiles above 128 channels -- one matrix per (width, tile))
        for L in (L for blk in arch.blocks for L in blk.layers if L.kind in ("res", "attn")):
            name = self.ident_name(L.cout, 9 if L.kind == "res" else 1)
            if name not in self.slots:
                self._reserve(name, 1, (1, L.cout, L.cout))
        K = self.unet.out_channels
        c_head = int(self.unet.channel_mult[0] * mc)
        self._reserve("out:g", c_head); self._reserve("out:be", c_head)
        self._reserve("out:w", 9 * c_head * _ceil(K, 32), (9, c_head, K)); self._reserve("out:b", _ceil(K, 32))

    def _nt(self, cout: int, taps: int) -> int:
        """N tile of the tensor-core conv with `taps` taps and `cout` output channels (the kernel's own rule)."""
        return int(_lib.lib().ccdm_conv_tc_nt(int(cout), int(taps), 1 if self.x3 else 0))

    def ident_name(self, c: int, taps: int) -> str:
        """Slot of the c x c identity matrix packed for the N tile of a conv with `taps` taps (its fused residual chunk)."""
        return "ident:%d:%d" % (c, self._nt(c, taps))

    def addr(self, name) -> int:
        return self.buf.data_ptr() + 4 * self.slots[name][0]

    def addr16(self, name) -> int:
        return self.buf16.data_ptr() + 2 * self.slots16[name][0]

    def view(self, name) -> torch.Tensor:
        off, n = self.slots[name]
        return self.buf[off:off + n]

    def _current_stamp(self):
        return tuple((p.data_ptr(), p._version) for p in self.unet.parameters())

    def invalidate(self):
        """Force a repack on the next call.  The staleness stamp is (data_ptr, _version) of every parameter: it sees
        ``load_state_dict``, ``.to()`` and in-place tensor ops (the reference's Polyak averager writes ``dst[...] = value``), but
        NOT writes through ``p.data`` (``p.data.mul_()``, ``p.data.copy_()``), which do not bump ``_version`` -- code that updates
        weights that way calls this (``model.unet.engine(prec).weights.invalidate()``)."""
        self._stamp = None

    def refresh(self) -> bool:
        """(Re)pack if any parameter changed since the last call.  Returns True if repacked."""
        stamp = self._current_stamp()
        if self.buf is not None and stamp == self._stamp:
            return False
        if self.buf is None:
            self.buf = torch.zeros(self.total, dtype=torch.float32, device=self.device)
            self.buf16 = torch.zeros(max(self.total16, 8), dtype=torch.float16 if self.x3 else torch.bfloat16, device=self.device)
        sd = {k: v.detach().to(device=self.device, dtype=torch.float32) for k, v in self.unet.state_dict().items()}
        if self.x3:
            # one power-of-two scale for every packed weight, as large as fp16 allows (|2^shift w| <= 2^15; the sub-pixel
            # sums of the upsampling convs add up to four taps): the lo parts stay normal numbers, the identity matrices of
            # the residual chunks are 2^shift exactly, and the epilogue's 2^-acc_shift is exact
            wmax = 1e-30
            for k, v in sd.items():
                if v.dim() >= 3:
                    wmax = max(wmax, float(v.abs().max()) * (4.0 if k.endswith(".conv.weight") else 1.0))
            self.shift = int(max(0, min(13, math.floor(math.log2(32768.0 / wmax)))))

        def put(name, t, raw=None, nt=None):
            # nt: N tile of the conv that reads these weights when it is not the tile of a stand-alone conv of their shape
            # (the 1x1 skip conv and the identity residual are extra K chunks of a 3x3 conv and follow ITS tile)
            v = self.view(name)
            v.zero_()
            v[:t.numel()].copy_(t.reshape(-1))
            if raw is not None and name in self.slots16:
                off, n = self.slots16[name]
                tc = (pack_conv_weight_x3(raw, self.shift, nt) if self.x3 else pack_conv_weight_tc(raw, nt)).reshape(-1)
                assert tc.numel() == n, (name, tc.numel(), n)
                self.buf16[off:off + n].copy_(tc)

        conv_w, padded = pack_conv_weight, pack_bias

        put("te0_w", sd["time_embed.0.weight"]); put("te0_b", sd["time_embed.0.bias"])
        put("te2_w", sd["time_embed.2.weight"]); put("te2_b", sd["time_embed.2.bias"])
        emb_w, emb_b = [], []
        for blk in self.unet.arch.blocks:
            for L in blk.layers:
                p = L.path
                if L.kind == "conv_in":
                    w_in = sd[p + ".weight"]
                    w_pad = torch.zeros((w_in.shape[0], _ceil(L.cin, 16)) + tuple(w_in.shape[2:]), dtype=w_in.dtype, device=w_in.device)
                    w_pad[:, :L.cin] = w_in
                    put(p + ":w", conv_w(w_in, _ceil(L.cin, 8)), w_pad); put(p + ":b", padded(sd[p + ".bias"]))
                elif L.kind in ("down", "up"):
                    raw = subpixel_weights(sd[p + ".weight"]) if L.kind == "up" else sd[p + ".weight"]
                    put(p + ":w", conv_w(sd[p + ".weight"]), raw); put(p + ":b", padded(sd[p + ".bias"]))
                elif L.kind == "res":
                    put(p + ":g1", sd[p + ".in_layers.0.weight"]); put(p + ":be1", sd[p + ".in_layers.0.bias"])
                    put(p + ":w1", conv_w(sd[p + ".in_layers.2.weight"]), sd[p + ".in_layers.2.weight"]); put(p + ":b1", sd[p + ".in_layers.2.bias"])
                    put(p + ":g2", sd[p + ".out_layers.0.weight"]); put(p + ":be2", sd[p + ".out_layers.0.bias"])
                    put(p + ":w2", conv_w(sd[p + ".out_layers.3.weight"]), sd[p + ".out_layers.3.weight"])
                    b2 = sd[p + ".out_layers.3.bias"]
                    if L.skip_conv:
                        ws = sd[p + ".skip_connection.weight"]  # [Cout, Cin, 1, 1] -> [Cin][Cout]
                        put(p + ":ws", ws[:, :, 0, 0].t().contiguous(), ws, nt=self._nt(L.cout, 9) if self.with_tc else None)
                        b2 = b2 + sd[p + ".skip_connection.bias"]
                    put(p + ":b2", b2)
                    if p + ":w1s" in self.slots:
                        f = blk.feat_concat
                        c_h = L.cin - f
                        _, n_f, f16 = feat_fold_split(c_h, f)
                        cs = c_h + f16
                        w1, g1, be1 = sd[p + ".in_layers.2.weight"], sd[p + ".in_layers.0.weight"], sd[p + ".in_layers.0.bias"]
                        w1s = w1[:, :cs].clone()
                        w1s[:, c_h + n_f:] = 0          # feature channels past the shared group belong to the per-chain map
                        w1m = w1[:, c_h:].clone()
                        w1m[:, :n_f] = 0                # ... and the shared group's feature channels to the per-step conv
                        put(p + ":g1s", g1[:cs]); put(p + ":be1s", be1[:cs]); put(p + ":w1s", conv_w(w1s), w1s)
                        put(p + ":g1m", g1[c_h:]); put(p + ":be1m", be1[c_h:]); put(p + ":w1m", conv_w(w1m), w1m)
                        ws = sd[p + ".skip_connection.weight"]
                        wsm = ws[:, c_h:].contiguous()
                        put(p + ":wsm", wsm[:, :, 0, 0].t().contiguous(), wsm)
                        ws2 = torch.cat([ws[:, :c_h], torch.eye(L.cout, device=ws.device).reshape(L.cout, L.cout, 1, 1)], 1).contiguous()
                        put(p + ":ws2", ws2[:, :, 0, 0].t().contiguous(), ws2, nt=self._nt(L.cout, 9) if self.with_tc else None)
                        put(p + ":b0", torch.zeros(1, device=self.device))
                    emb_w.append(sd[p + ".emb_layers.1.weight"]); emb_b.append(sd[p + ".emb_layers.1.bias"])
                elif L.kind == "attn":
                    put(p + ":g", sd[p + ".norm.weight"]); put(p + ":be", sd[p + ".norm.bias"])
                    put(p + ":wqkv", sd[p + ".qkv.weight"][:, :, 0].t().contiguous(), sd[p + ".qkv.weight"]); put(p + ":bqkv", sd[p + ".qkv.bias"])
                    put(p + ":wproj", sd[p + ".proj_out.weight"][:, :, 0].t().contiguous(), sd[p + ".proj_out.weight"]); put(p + ":bproj", sd[p + ".proj_out.bias"])
        put("emb_w", torch.cat(emb_w, 0)); put("emb_b", torch.cat(emb_b, 0))
        for name in self.slots16:
            if name.startswith("ident:"):
                c, nt = (int(v) for v in name.split(":")[1:])
                put(name, torch.zeros(1, device=self.device), torch.eye(c, device=self.device).reshape(c, c, 1, 1), nt=nt)
        put("out:g", sd["out.0.weight"]); put("out:be", sd["out.0.bias"])
        put("out:w", conv_w(sd["out.2.weight"]), sd["out.2.weight"]); put("out:b", padded(sd["out.2.bias"]))
        self._stamp = stamp
        return True


# --------------------------------------------------------------------------------------
# program
# --------------------------------------------------------------------------------------
class Program:
    """One reverse step for fixed (B, H, W): op list + workspace, bound to device addresses."""

    def __init__(self, engine: "UNetEngine", B: int, H: int, W: int, rows_per_sample: int, img_rep: int = 1, tile_batch: int = 0):
        unet, arch = engine.unet, engine.unet.arch
        self.engine = engine
        self.B, self.H, self.W = B, H, W
        if img_rep < 1 or B % img_rep:
            raise ValueError(f"batch {B} is not a multiple of the samples per image ({img_rep})")
        self.img_rep = img_rep  # consecutive samples that share one conditioning image / feature map
        self.tile_batch = tile_batch  # ccdm_op::tile_batch of every conv: batch-independent tiling (0: tiles follow B)
        self.K = unet.out_channels
        self.C_img = unet.in_channels - self.K
        self.dt, self.esize, tc_mode = PRECISIONS[engine.precision]
        self.exact = 0 if tc_mode else 1     # ccdm_op::exact of the convs / attention: 1 = FFMA kernels
        self.x3 = engine.precision == "exact"
        dev = engine.device
        L = _lib.lib()

        ops: List[dict] = []      # op field dicts with Ten references, resolved after allocation
        tens: List[Ten] = []

        def new(name, C, h, w, stat=True, esize=None):
            t = Ten(name, C, h, w, esize or self.esize, stat, fmt=_lib.DT_F32 if esize == 4 else self.dt)
            t.nbytes = _ceil(B * h * w * C * t.esize, ALIGN)
            tens.append(t)
            return t

        def emit(kind, ins, out, **f):
            i = len(ops)
            for t in ins:
                if t is not None:
                    t.last = max(t.last, i)
            if out is not None:
                out.born = i
                out.last = max(out.last, i)
            f.update(kind=kind, _ins=ins, _out=out)
            ops.append(f)
            return out

        self.feat: Optional[Ten] = None
        fc = unet.feat_channels
        if fc:
            self.feat = Ten("feature_condition", fc, H // 8, W // 8, self.esize, True, fmt=self.dt, external=True)

        # per-chain constants of the folded feature-concat ResBlock: set where that block is planned
        self.feat16: Optional[Ten] = None   # the leading feature channels that share a GroupNorm group with h
        self.fmaps: List[Ten] = []          # [conv map, skip map]
        self._pre_dicts: List[dict] = []    # the ops that produce them, launched once per chain (run_pre)
        hs: List[Ten] = []
        h: Optional[Ten] = None
        ch, cw = H, W
        for blk in arch.blocks:
            srcs: List[Ten] = [h] if h is not None else []
            if blk.feat_concat:
                if self.feat is None or (self.feat.H, self.feat.W) != (ch, cw):
                    raise ValueError("feature condition resolution does not match its target layer")
                srcs = [h, self.feat]
            if blk.stage == "out":
                srcs = [h, hs.pop()]
            for Ly in blk.layers:
                p = Ly.path
                if Ly.kind == "conv_in" and self.exact == 0:
                    # bf16: materialise one-hot(x_t) ++ image as a plane-major tensor (16 bytes per pixel and plane),
                    # then input_blocks[0] is an ordinary tensor-core conv without a norm
                    cp = _ceil(Ly.cin, 16)
                    xin = emit(_lib.OP_ENCODE_INPUT, [], new(p + ":x", cp, ch, cw, stat=False), Hin=ch, Win=cw, Hout=ch, Wout=cw,
                               Cout=cp, K=self.K, C_img=self.C_img, img_rep=img_rep, _labels=True)
                    h = emit(_lib.OP_CONV, [xin], new(p, Ly.cout, ch, cw), ksize=3, stride=1, gn=0, silu=0, Hin=ch, Win=cw,
                             Hout=ch, Wout=cw, Cout=Ly.cout, _src=[xin], _w=p + ":w", _b=p + ":b")
                elif Ly.kind == "conv_in":
                    h = emit(_lib.OP_INPUT_CONV, [], new(p, Ly.cout, ch, cw), src_kind=1, ksize=3, stride=1, Hin=ch, Win=cw,
                             Hout=ch, Wout=cw, Cout=Ly.cout, K=self.K, C_img=self.C_img, img_rep=img_rep, _w=p + ":w", _b=p + ":b")
                elif (Ly.kind == "res" and engine.fold_features and blk.feat_concat and len(srcs) == 2 and srcs[1] is self.feat and Ly.skip_conv
                      and (p + ":w1s") in engine.weights.slots and not self.fmaps):
                    # input_blocks[10] with the constant feature channels folded into two per-chain maps:
                    #   h1 = conv3x3(SiLU(GN(cat[h, f])))  =  conv over cat[h, f[:f16]] (the groups h takes part in; per step)
                    #                                         + fmap1 = conv over f with the same 14-channel groups (per chain)
                    #   skip(cat[h, f])                    =  skip over h (per step) + fmap2 = skip over f (per chain)
                    hh, ft = srcs
                    cpg, n_f, f16 = feat_fold_split(hh.C, ft.C)
                    if f16:
                        self.feat16 = Ten("feature_condition[:%d]" % f16, f16, ch, cw, self.esize, True, fmt=self.dt, external=True)
                    self.fmaps = [Ten(p + ":fmap%d" % j, Ly.cout, ch, cw, self.esize, False, fmt=self.dt, external=True) for j in (1, 2)]
                    geo = dict(stride=1, Hin=ch, Win=cw, Hout=ch, Wout=cw, Cout=Ly.cout)
                    self._pre_dicts = [
                        dict(kind=_lib.OP_CONV, ksize=3, gn=1, silu=1, gn_cpg=cpg, gn_off=hh.C, _src=[ft], _g=p + ":g1m", _be=p + ":be1m",
                             _w=p + ":w1m", _b=p + ":b0", _ins=[ft], _out=self.fmaps[0], **geo),
                        dict(kind=_lib.OP_CONV, ksize=1, gn=0, silu=0, _src=[ft], _w=p + ":wsm", _b=p + ":b0", _ins=[ft], _out=self.fmaps[1], **geo)]
                    s1 = [hh] + ([self.feat16] if f16 else [])
                    h1 = emit(_lib.OP_CONV, s1, new(p + ":h1", Ly.cout, ch, cw), ksize=3, gn=1, silu=1, gn_cpg=cpg, gn_off=0, _src=s1,
                              _g=p + ":g1s", _be=p + ":be1s", _w=p + ":w1s", _b=p + ":b1", emb_off=Ly.emb_off, _emb=True,
                              _skip=[self.fmaps[0]], _ws=engine.weights.ident_name(Ly.cout, 9), **geo)
                    h = emit(_lib.OP_CONV, [h1, hh], new(p, Ly.cout, ch, cw), ksize=3, gn=1, silu=1, _src=[h1], _g=p + ":g2", _be=p + ":be2",
                             _w=p + ":w2", _b=p + ":b2", _skip=[hh, self.fmaps[1]], _ws=p + ":ws2", **geo)
                    srcs = [h]
                elif Ly.kind == "res":
                    if len(srcs) == 2 and not Ly.skip_conv:
                        raise NotImplementedError("identity skip over a concatenated input")
                    if sum(s.C for s in srcs) != Ly.cin:
                        raise ValueError(f"{p}: expected {Ly.cin} input channels, got {[s.C for s in srcs]}")
                    h1 = emit(_lib.OP_CONV, srcs, new(p + ":h1", Ly.cout, ch, cw), ksize=3, stride=1, gn=1, silu=1, Hin=ch,
                              Win=cw, Hout=ch, Wout=cw, Cout=Ly.cout, _src=srcs, _g=p + ":g1", _be=p + ":be1", _w=p + ":w1",
                              _b=p + ":b1", emb_off=Ly.emb_off, _emb=True)
                    f = dict(ksize=3, stride=1, gn=1, silu=1, Hin=ch, Win=cw, Hout=ch, Wout=cw, Cout=Ly.cout, _src=[h1],
                             _g=p + ":g2", _be=p + ":be2", _w=p + ":w2", _b=p + ":b2")
                    if Ly.skip_conv:
                        f.update(_skip=srcs, _ws=p + ":ws")
                    elif self.exact == 0 and (_IDENT_SKIP or self.x3):
                        f.update(_skip=[srcs[0]], _ws=engine.weights.ident_name(Ly.cout, 9))
                    else:
                        f.update(_res=srcs[0])
                    h = emit(_lib.OP_CONV, [h1] + srcs, new(p, Ly.cout, ch, cw), **f)
                    srcs = [h]
                elif Ly.kind == "attn":
                    x = srcs[0]
                    C = Ly.cin
                    qkv = emit(_lib.OP_CONV, [x], new(p + ":qkv", 3 * C, ch, cw, stat=False), ksize=1, stride=1, gn=1, silu=0,
                               Hin=ch, Win=cw, Hout=ch, Wout=cw, Cout=3 * C, _src=[x], _g=p + ":g", _be=p + ":be",
                               _w=p + ":wqkv", _b=p + ":bqkv")
                    hd = C // Ly.heads
                    if hd * Ly.heads != C or hd not in (32, 64):
                        # (every shipped configuration has num_head_channels = 32, params_eval.yml:56-63)
                        raise NotImplementedError(f"{p}: attention head_dim {C}/{Ly.heads} is not implemented by the B200 sampler (32 or 64)")
                    a = emit(_lib.OP_ATTENTION, [qkv], new(p + ":a", C, ch, cw, stat=False), Hin=ch, Win=cw, Hout=ch, Wout=cw,
                             Cout=C, heads=Ly.heads, head_dim=C // Ly.heads, _src=[qkv])
                    fr = dict(_skip=[x], _ws=engine.weights.ident_name(C, 1)) if self.exact == 0 and (_IDENT_SKIP or self.x3) else dict(_res=x)
                    h = emit(_lib.OP_CONV, [a, x], new(p, C, ch, cw), ksize=1, stride=1, gn=0, silu=0, Hin=ch, Win=cw, Hout=ch,
                             Wout=cw, Cout=C, _src=[a], _w=p + ":wproj", _b=p + ":bproj", **fr)
                    srcs = [h]
                elif Ly.kind == "down":
                    nh, nw = (ch + 1) // 2, (cw + 1) // 2
                    h = emit(_lib.OP_CONV, srcs, new(p, Ly.cout, nh, nw), ksize=3, stride=2, Hin=ch, Win=cw, Hout=nh, Wout=nw,
                             Cout=Ly.cout, _src=srcs, _w=p + ":w", _b=p + ":b")
                    ch, cw = nh, nw
                    srcs = [h]
                elif Ly.kind == "up":
                    h = emit(_lib.OP_CONV, srcs, new(p, Ly.cout, ch * 2, cw * 2), ksize=3, stride=1, upsample=1, Hin=ch, Win=cw,
                             Hout=ch * 2, Wout=cw * 2, Cout=Ly.cout, _src=srcs, _w=p + ":w", _b=p + ":b")
                    ch, cw = ch * 2, cw * 2
                    srcs = [h]
            if blk.stage == "in":
                hs.append(h)
                h.last = 10 ** 9  # provisional: popped later, fixed when consumed
        assert (ch, cw) == (H, W) and not hs
        logits = emit(_lib.OP_CONV, [h], new("logits", self.K, H, W, stat=False, esize=4), ksize=3, stride=1, gn=1, silu=1,
                      Hin=H, Win=W, Hout=H, Wout=W, Cout=self.K, _src=[h], _g="out:g", _be="out:be", _w="out:w", _b="out:b",
                      out_dtype=_lib.DT_F32)
        emit(_lib.OP_HEAD, [logits], None, Hin=H, Win=W, K=self.K, _src=[logits])

        # skip tensors: recompute true last use (emit() already recorded every reader)
        for t in tens:
            if t.last >= 10 ** 9:
                t.last = max(i for i, o in enumerate(ops) if t in o["_ins"])

        # ---- allocate ---------------------------------------------------------------
        arena = Arena()
        stat_bytes = 0
        for i, o in enumerate(ops):
            out = o["_out"]
            if out is not None:
                arena.alloc(out)
                if out.want_stat:
                    out.stat_off = stat_bytes
                    stat_bytes += _ceil(B * out.C * 16, ALIGN)
            arena.release_dead(i)
        part_floats = 1  # sized in bind(), once each op's kernel (FFMA / tcgen05) is known
        n_pix = B * H * W
        feat_bytes = _ceil(B * fc * (H // 8) * (W // 8) * self.esize, ALIGN) if fc else 0
        layout = dict(arena=arena.size, stat=stat_bytes, part=_ceil(part_floats * 4, ALIGN), ticket=_ceil(4 * (B + 1), ALIGN),
                      labels=_ceil(n_pix, ALIGN), image=_ceil(n_pix * self.C_img * 4, ALIGN), feat=feat_bytes,
                      feat_stat=_ceil(B * fc * 16, ALIGN) if fc else 0, probs=_ceil(n_pix * self.K * 4, ALIGN),
                      noise=_ceil(n_pix * self.K * 4, ALIGN), step=ALIGN)
        if self.fmaps:
            fh, fw = self.fmaps[0].H, self.fmaps[0].W
            layout.update(fmap1=_ceil(B * self.fmaps[0].C * fh * fw * self.esize, ALIGN), fmap2=_ceil(B * self.fmaps[1].C * fh * fw * self.esize, ALIGN))
            if self.feat16 is not None:
                layout.update(feat16=_ceil(B * self.feat16.C * fh * fw * self.esize, ALIGN), feat16_stat=_ceil(B * self.feat16.C * 16, ALIGN))
        offs, total = {}, 0
        for k, v in layout.items():
            offs[k] = total
            total += v
        self.workspace = torch.zeros(total, dtype=torch.uint8, device=dev)
        base = self.workspace.data_ptr()
        self.addr = {k: base + v for k, v in offs.items()}
        self.nbytes = total
        self.arena_bytes = arena.size
        if self.feat is not None:
            self.feat.addr = self.addr["feat"]
            self.feat.stat_addr = self.addr["feat_stat"]
        if self.fmaps:
            self.fmaps[0].addr, self.fmaps[1].addr = self.addr["fmap1"], self.addr["fmap2"]
            if self.feat16 is not None:
                self.feat16.addr, self.feat16.stat_addr = self.addr["feat16"], self.addr["feat16_stat"]
        for t in tens:
            t.addr = self.addr["arena"] + t.off
            t.stat_addr = self.addr["stat"] + t.stat_off if t.want_stat else 0

        # typed views for the host side
        def view(key, dtype, shape):
            n = int(torch.tensor([], dtype=dtype).element_size()) * math.prod(shape)
            return self.workspace[offs[key]:offs[key] + n].view(dtype).view(*shape)

        self.labels = view("labels", torch.uint8, (B, H, W))
        self.image = view("image", torch.float32, (B, self.C_img, H, W))
        self.probs = view("probs", torch.float32, (B, H, W, self.K))
        self.noise = view("noise", torch.float32, (n_pix, self.K))
        self.step_counter = view("step", torch.int32, (1,))
        self.logits = None
        self.tens = {t.name: t for t in tens}

        # step table + embedding table are (re)bound per run
        self.max_rows = 0
        self.steps_buf: Optional[torch.Tensor] = None
        self.emb_buf: Optional[torch.Tensor] = None
        self.rows_per_sample = rows_per_sample
        self._op_dicts = ops
        self.plan = None
        self.n_ops = len(ops)
        self._bound_rows = -1

    # -- binding ---------------------------------------------------------------------
    def bind(self, n_rows: int):
        """(Re)build the C plan against step/emb tables of at least ``n_rows`` rows."""
        L = _lib.lib()
        if self.plan is not None and n_rows <= self.max_rows:
            return
        if self.plan is not None:
            L.ccdm_plan_destroy(self.plan)
            self.plan = None
        eng = self.engine
        dev = eng.device
        self.max_rows = max(n_rows, 8)
        emb_cols = eng.unet.arch.emb_cols
        self.steps_buf = torch.zeros(self.max_rows * ctypes.sizeof(StepEntry), dtype=torch.uint8, device=dev)
        self.emb_buf = torch.zeros(self.max_rows * max(emb_cols, 1), dtype=torch.float32, device=dev)
        self.t_buf = torch.zeros(self.max_rows, dtype=torch.float32, device=dev)
        W = eng.weights
        self._stat_bufs = []
        for t in self.tens.values():
            t.part_addr, t.stat_layout = 0, None

        def base_op(o) -> Op:
            """Shape / variant fields of an op: everything the kernel dispatch depends on (no statistics, no weights)."""
            f = {k: v for k, v in o.items() if not k.startswith("_")}
            fields = dict(dtype=self.dt, B=self.B, exact=self.exact, out_dtype=self.dt)
            if o["kind"] == _lib.OP_CONV:
                fields["tile_batch"] = self.tile_batch
            fields.update(f)
            op = Op(**fields)
            src = o.get("_src", [])
            if o["kind"] in (_lib.OP_CONV, _lib.OP_ATTENTION):
                op.src0, op.C0 = src[0].addr, src[0].C
                if len(src) > 1:
                    op.src1, op.C1 = src[1].addr, src[1].C
            if "_skip" in o:
                sk = o["_skip"]
                op.skip0, op.S0 = sk[0].addr, sk[0].C
                if len(sk) > 1:
                    op.skip1, op.S1 = sk[1].addr, sk[1].C
            if "_res" in o:
                op.res = o["_res"].addr
            return op

        # pass 1: which convs land on the tensor-core kernel.  Everything downstream -- weight layout, statistics layout of
        # the INPUT tensors -- follows from this, so it is decided before anything is bound.
        all_dicts = self._op_dicts + self._pre_dicts  # the per-chain ops (run_pre) are bound like the step's, behind them
        use_tc = [bool(self.exact == 0 and o["kind"] == _lib.OP_CONV and L.ccdm_conv_uses_tc(ctypes.byref(base_op(o))))
                  for o in all_dicts]
        off_tc = [i for i, o in enumerate(all_dicts) if self.exact == 0 and o["kind"] == _lib.OP_CONV and not use_tc[i]]
        if off_tc:
            what = ", ".join("op %d (%s -> %d ch @%dx%d)" % (i, "+".join(str(t.C) for t in all_dicts[i]["_src"]),
                                                              all_dicts[i]["Cout"], all_dicts[i]["Hout"], all_dicts[i]["Wout"])
                             for i in off_tc)
            if self.x3:
                raise _lib.CcdmError("precision='exact': the tensor-core conv kernel cannot take " + what +
                                     "; this mode has no other conv kernel (use precision='fp32')")
            LOGGER.warning("precision='bf16': %d conv(s) do not fit the tensor-core kernel and run on the fp32 FFMA kernel: %s",
                           len(off_tc), what)
        self.off_tc = off_tc
        # a tensor's statistics are left as per-CTA partial rows (deferred fold) only if its producer is a tensor-core conv
        # AND every GroupNorm consumer is one too: the FFMA kernel reads folded double2 sums
        gn_readers: Dict[int, List[int]] = {}
        for i, o in enumerate(all_dicts):
            if o["kind"] == _lib.OP_CONV and o.get("gn"):
                for sten in o.get("_src", [])[:2]:
                    gn_readers.setdefault(id(sten), []).append(i)

        arr: List[Op] = [None] * len(all_dicts)
        for i, o in enumerate(all_dicts):
            op = base_op(o)
            src = o.get("_src", [])
            if o["kind"] in (_lib.OP_INPUT_CONV, _lib.OP_ENCODE_INPUT):
                op.labels_in = self.addr["labels"]
                op.image = self.addr["image"]
            if o["kind"] == _lib.OP_CONV and o.get("gn"):
                for si, sten in enumerate(src[:2]):
                    if sten.stat_layout is not None:  # the producer left per-CTA partial rows: this op folds them
                        assert use_tc[i]
                        setattr(op, "stat%d" % si, sten.part_addr)
                        for name, v in zip(("st_slots", "st_ips", "st_items", "st_grid", "st_rows"), sten.stat_layout):
                            setattr(op, "%s%d" % (name, si), int(v))
                    else:
                        setattr(op, "stat%d" % si, sten.stat_addr)
            if "_g" in o:
                op.gamma, op.beta = W.addr(o["_g"]), W.addr(o["_be"])
            # attention: exact=0 selects the tcgen05 kernel (head_dim 32); head: exact=0 selects fast maths for sampling steps
            if o["kind"] == _lib.OP_CONV and not use_tc[i]:
                op.exact = 1
            if o["kind"] == _lib.OP_HEAD and self.x3:
                op.exact = 1  # the exact tensor-core mode draws with the bit-exact posterior arithmetic
            if use_tc[i] and self.x3:
                op.acc_shift = W.shift + _lib.F16X2_SCALE_LOG2
            if "_w" in o:
                op.weight, op.bias = (W.addr16(o["_w"]) if use_tc[i] else W.addr(o["_w"])), W.addr(o["_b"])
            if o.get("_emb"):
                op.emb = self.emb_buf.data_ptr()
                op.emb_cols = emb_cols
                op.emb_bstride = self.rows_per_sample
            if "_skip" in o:
                if not use_tc[i] and o["_ws"].startswith("ident:"):
                    # identity residual folded into the MMA exists as packed tensor-core weights only: on the FFMA kernel
                    # the residual is read in the epilogue instead
                    op.skip0 = op.skip1 = op.S0 = op.S1 = 0
                    op.res = o["_skip"][0].addr
                else:
                    op.skip_w = W.addr16(o["_ws"]) if use_tc[i] else W.addr(o["_ws"])
            out = o["_out"]
            if out is not None:
                op.out = out.addr
                deferred = use_tc[i] and all(use_tc[j] for j in gn_readers.get(id(out), []))
                if out.want_stat and deferred:
                    # deferred fold: the epilogue only writes its per-CTA partial rows into a buffer owned by the
                    # tensor; the consumers' GroupNorm prologue folds them (no ticket, fence or atomic in the producer)
                    lay = (ctypes.c_int32 * 5)()
                    _lib.check(L.ccdm_conv_stat_layout(ctypes.byref(op), lay), "conv_stat_layout")
                    op.part = 1  # so that ccdm_op_part_floats sees a statistics-producing op
                    nfl = int(L.ccdm_op_part_floats(ctypes.byref(op)))
                    buf = torch.zeros(max(nfl, 2), dtype=torch.float32, device=dev)
                    self._stat_bufs.append(buf)
                    out.part_addr, out.stat_layout = buf.data_ptr(), tuple(int(v) for v in lay)
                    op.part, op.ostat, op.ticket = out.part_addr, 0, 0
                elif out.want_stat:
                    op.ostat = out.stat_addr
                    op.part = 1  # patched below once the scratch size is known
                    op.ticket = self.addr["ticket"]
            if o["kind"] == _lib.OP_HEAD:
                op.src0 = src[0].addr
                op.labels_in = self.addr["labels"]
                op.labels_out = self.addr["labels"]
                op.probs_out = self.addr["probs"]
                op.noise = self.addr["noise"]
                op.ticket = self.addr["ticket"] + 4 * self.B
            op.steps = self.steps_buf.data_ptr()
            op.step_ptr = self.addr["step"]
            arr[i] = op
        shared = [i for i in range(len(arr)) if arr[i].part == 1]  # folded-by-producer ops share one scratch
        part_floats = max([int(L.ccdm_op_part_floats(ctypes.byref(arr[i]))) for i in shared] + [1])
        self.part_buf = torch.zeros(part_floats, dtype=torch.float32, device=dev)
        for i in shared:
            arr[i].part = self.part_buf.data_ptr()
        self.n_tc = sum(use_tc[:self.n_ops])
        self._op_array = (Op * self.n_ops)(*arr[:self.n_ops])
        self._pre_array = arr[self.n_ops:]
        self.plan = L.ccdm_plan_create(self._op_array, self.n_ops)
        if not self.plan:
            raise _lib.CcdmError("ccdm_plan_create failed: " + L.ccdm_last_error().decode())

    def run_pre(self, stream_ptr):
        """Launch the per-chain ops (the constant feature maps of the folded feature-concat ResBlock).  Call after bind() and
        after the feature condition of the chain has been loaded."""
        for op in self._pre_array:
            _lib.check(_lib.lib().ccdm_launch_op(ctypes.byref(op), stream_ptr), "per-chain feature map")

    def set_noise(self, noise_mode: int, seed: int = 0, sample0: int = 0, use_noise_buf: bool = False,
                  export_noise: bool = False):
        """Per-run fields of the head op, on the C plan and on the host mirror used for single launches."""
        L = _lib.lib()
        nz = self.noise.data_ptr() if use_noise_buf else 0
        nz_out = self.noise.data_ptr() if export_noise else 0
        _lib.check(L.ccdm_plan_set_noise(self.plan, noise_mode, int(seed), int(sample0), ctypes.c_void_p(nz or None),
                                         ctypes.c_void_p(nz_out or None)), "plan_set_noise")
        for i in range(self.n_ops):
            if self._op_array[i].kind == _lib.OP_HEAD:
                o = self._op_array[i]
                o.noise_mode, o.seed, o.sample0, o.noise, o.noise_out = noise_mode, int(seed), int(sample0), nz, nz_out

    def __del__(self):
        try:
            if self.plan is not None:
                _lib.lib().ccdm_plan_destroy(self.plan)
        except Exception:
            pass


# --------------------------------------------------------------------------------------
# engine
# --------------------------------------------------------------------------------------
class UNetEngine:
    def __init__(self, unet, precision: str = "fp32", dry_run: bool = False):
        """``dry_run``: plan programs with host buffers and never launch (host-logic tests
        on machines without a GPU); any attempt to run raises."""
        if precision not in PRECISIONS:
            raise ValueError("precision must be one of %s" % sorted(PRECISIONS))
        p0 = next(unet.parameters())
        self.dry_run = dry_run
        if not dry_run:
            _lib.require_device()
            if p0.device.type != "cuda":
                raise _lib.CcdmError("the model must live on a CUDA device (B200); there is no CPU path")
        self.unet = unet
        self.precision = precision
        self.device = p0.device
        self.weights = PackedWeights(unet, self.device, with_tc=PRECISIONS[precision][2], x3=(precision == "exact"))
        self.programs: "OrderedDict[tuple, Program]" = OrderedDict()
        self.stream = None if dry_run else torch.cuda.Stream(device=self.device)
        self.use_graph = True
        # Lanes (experimental, default 1): the batch of a chain is split into `lanes` contiguous sub-batches that run as
        # independent programs on their own streams, in the hope that the latency chains of the ~50 small launches per
        # step overlap across sub-batches.  Philox noise is keyed by the global sample index, so the result does not
        # depend on the split (tested).  MEASURED: no overlap happens -- LIDC B=64 2.37 / 2.84 / 4.14 ms per step at
        # 1 / 2 / 4 lanes (the sub-batch steps serialise; every kernel asks for most of an SM's shared memory) -- so
        # it stays off; see DESIGN.md section 9.
        import os
        self.lanes = max(1, int(os.environ.get("CCDM_LANES", "1")))
        # Batch-independent tiling (ccdm_op::tile_batch): 0 = every conv picks its tile height from its own batch (fastest);
        # N > 0 = as if the batch were N.  With a fixed value the `exact` mode's result for a sample is bit-identical
        # whatever batch it runs in (DenoisingModel.tile_batch; sample_sharded sets 64).
        self.tile_batch = 0
        # run the feature-concat ResBlock in its folded form (constant DINO channels as per-chain maps); False keeps
        # the per-step 448-channel concatenation (A/B, tests)
        self.fold_features = _FOLD_FEAT
        self._children: List["UNetEngine"] = []

    # -- helpers ----------------------------------------------------------------------
    def program(self, B, H, W, rows_per_sample=0, img_rep=1) -> Program:
        key = (B, H, W, rows_per_sample, img_rep, self.tile_batch, self.fold_features)
        prog = self.programs.get(key)
        if prog is None:
            # bounded cache: a ragged last batch or a change of batch size must not pile up ~1 GB workspaces
            while len(self.programs) >= PROGRAM_CACHE:
                self.programs.popitem(last=False)
            prog = self.programs[key] = Program(self, B, H, W, rows_per_sample, img_rep, self.tile_batch)
        else:
            self.programs.move_to_end(key)
        return prog

    def _sp(self):
        return _lib.stream_ptr(self.stream)

    def _load_inputs(self, prog: Program, x, condition, feature_condition):
        L = _lib.lib()
        dev = self.device
        B, H, W, K = prog.B, prog.H, prog.W, prog.K
        if x.dtype == torch.uint8 and x.dim() == 3:
            prog.labels.copy_(x.to(dev, non_blocking=True))
        else:
            if tuple(x.shape) != (B, K, H, W):
                raise ValueError(f"x must be one-hot [B,{K},H,W] (or uint8 labels [B,H,W]); got {tuple(x.shape)}")
            xf = x.to(device=dev, dtype=torch.float32, non_blocking=True)
            sb, sk, sh, sw = xf.stride()
            _lib.check(L.ccdm_onehot_to_labels(xf.data_ptr(), sb, sk, sh, sw, B, K, H, W, prog.labels.data_ptr(), self._sp()),
                       "onehot_to_labels")
            xf.record_stream(self.stream)
        n_img = B // prog.img_rep  # img_rep samples share one image: the kernels index it by sample // img_rep
        if tuple(condition.shape) != (n_img, prog.C_img, H, W):
            raise ValueError(f"condition must be [{n_img},{prog.C_img},{H},{W}]; got {tuple(condition.shape)}")
        prog.image[:n_img].copy_(condition.to(device=dev, dtype=torch.float32, non_blocking=True))
        if prog.feat is not None:
            if feature_condition is None:
                raise ValueError("this UNet was built with a feature_cond_encoder: feature_condition is required")
            f = prog.feat
            if tuple(feature_condition.shape) != (n_img, f.C, f.H, f.W):
                raise ValueError(f"feature_condition must be [{n_img},{f.C},{f.H},{f.W}]; got {tuple(feature_condition.shape)}")
            fcond = feature_condition.to(device=dev, dtype=torch.float32, non_blocking=True).contiguous()
            _lib.check(L.ccdm_nchw_to_nhwc_stats(fcond.data_ptr(), B, f.C, f.H, f.W, prog.dt, f.addr, f.stat_addr, prog.img_rep,
                                                 self._sp()), "nchw_to_nhwc_stats")
            if prog.feat16 is not None:
                f16 = prog.feat16
                head = fcond[:, :f16.C].contiguous()
                _lib.check(L.ccdm_nchw_to_nhwc_stats(head.data_ptr(), B, f16.C, f16.H, f16.W, prog.dt, f16.addr, f16.stat_addr, prog.img_rep,
                                                     self._sp()), "nchw_to_nhwc_stats (leading feature channels)")
                head.record_stream(self.stream)
            fcond.record_stream(self.stream)
        # (unet.py:770: a feature_condition passed to a UNet without an encoder slot is ignored)

    def _write_tables(self, prog: Program, entries: Sequence[tuple], t_rows: Sequence[float]):
        """entries: (t, alpha, cumalpha_tm1, mode, draw, emb_row) per step; t_rows: timesteps of the embedding rows."""
        L = _lib.lib()
        n_rows = max(len(entries), len(t_rows))
        prog.bind(n_rows)
        arr = (StepEntry * len(entries))()
        for i, (t, a, c, mode, draw, row) in enumerate(entries):
            arr[i] = StepEntry(float(t), float(a), float(c), int(mode), int(draw), int(row), 0, 0)
        host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
        prog.steps_buf[:host.numel()].copy_(host, non_blocking=False)
        prog.t_buf[:len(t_rows)].copy_(torch.tensor(list(t_rows), dtype=torch.float32))
        prog.step_counter.zero_()
        Wt = self.weights
        mc = self.unet.model_channels
        cols = self.unet.arch.emb_cols
        _lib.check(L.ccdm_time_table(prog.t_buf.data_ptr(), len(t_rows), mc, Wt.addr("te0_w"), Wt.addr("te0_b"), Wt.addr("te2_w"),
                                     Wt.addr("te2_b"), Wt.addr("emb_w"), Wt.addr("emb_b"), cols, prog.emb_buf.data_ptr(), self._sp()),
                   "time_table")

    # -- reference UNetModel.forward (one evaluation) ------------------------------------
    @torch.no_grad()
    def single_step(self, x, condition, feature_condition, timesteps, softmax=True):
        if not softmax:
            raise NotImplementedError("softmax_output=False is not implemented by the B200 sampler")
        L = _lib.lib()
        if self.dry_run:
            raise _lib.CcdmError("dry-run engine cannot execute")
        B, _, H, W = x.shape if x.dim() == 4 else (x.shape[0], None, x.shape[1], x.shape[2])
        ts = timesteps.detach().float().reshape(-1).cpu().tolist()
        if len(ts) != B:
            raise ValueError("timesteps must have one entry per sample")
        if x.dim() == 4 and x.is_floating_point():
            xs = x.detach()
            if not bool(((xs.max(dim=1).values == 1) & (xs.sum(dim=1) == 1)).all()):
                raise ValueError("UNetModel.forward: x must be a one-hot label map (the hot path keeps x_t as labels)")
        cur = torch.cuda.current_stream(self.device)
        self.stream.wait_stream(cur)
        with torch.cuda.stream(self.stream):
            self.weights.refresh()
            prog = self.program(B, H, W, rows_per_sample=1)
            self._load_inputs(prog, x, condition, feature_condition)
            self._write_tables(prog, [(ts[0], 0.0, 1.0, _lib.DRAW_X0, 0, 0)], ts)
            prog.run_pre(self._sp())
            prog.set_noise(_lib.NOISE_PHILOX)
            _lib.check(L.ccdm_plan_step(prog.plan, 0, self._sp()), "plan_step")
            out = prog.probs.clone()
        cur.wait_stream(self.stream)
        return out.permute(0, 3, 1, 2)

    # -- reference DenoisingModel.forward_denoising -----------------------------------------
    @torch.no_grad()
    def _lane_engines(self, n: int) -> List["UNetEngine"]:
        while len(self._children) < n:
            c = UNetEngine.__new__(UNetEngine)
            c.dry_run, c.unet, c.precision, c.device, c.weights = False, self.unet, self.precision, self.device, self.weights
            c.programs, c.stream, c.use_graph, c.lanes, c._children = OrderedDict(), torch.cuda.Stream(device=self.device), self.use_graph, 1, []
            c.tile_batch = self.tile_batch
            c.fold_features = self.fold_features
            self._children.append(c)
        for c in self._children:
            c.use_graph = self.use_graph
            c.tile_batch = self.tile_batch
        return self._children[:n]

    @torch.no_grad()
    def _run_chain_lanes(self, n_lanes, x, condition, feature_condition, t_values, alphas, cumalphas, last_mode, seed, sample0):
        B = x.shape[0]
        cur = torch.cuda.current_stream(self.device)
        self.stream.wait_stream(cur)
        with torch.cuda.stream(self.stream):
            self.weights.refresh()
        base, extra = divmod(B, n_lanes)
        outs, b0 = [], 0
        for j, child in enumerate(self._lane_engines(n_lanes)):
            b1 = b0 + base + (1 if j < extra else 0)
            fc = feature_condition[b0:b1] if feature_condition is not None else None
            outs.append(child.run_chain(x[b0:b1], condition[b0:b1], fc, t_values, alphas, cumalphas, last_mode, noise="philox",
                                        seed=seed, sample0=sample0 + b0, _parent_stream=self.stream))
            b0 = b1
        for child in self._children[:n_lanes]:
            cur.wait_stream(child.stream)
        for lab, pr in outs:
            lab.record_stream(cur)
            if pr is not None:
                pr.record_stream(cur)
        labels = torch.cat([o[0] for o in outs], 0)
        probs = torch.cat([o[1] for o in outs], 0) if outs[0][1] is not None else None
        return labels, probs

    @torch.no_grad()
    def run_chain(self, x, condition, feature_condition, t_values: Sequence[int], alphas, cumalphas, last_mode: int,
                  noise: str = "torch", seed: int = 0, sample0: int = 0, record=None, _parent_stream=None, img_rep: int = 1):
        """Runs the reverse chain; returns (labels uint8 [B,H,W], probs fp32 [B,H,W,K] or None).

        ``alphas``/``cumalphas``: host float lists (the DiffusionModel buffers).  ``noise``:
        'torch' draws ``E = empty(B*H*W, K).exponential_(1)`` from the device's global
        generator once per step with t > 1, exactly the stream the reference's
        ``torch.multinomial`` consumes; 'philox' draws in-kernel (sharding-invariant).
        A sequence of tensors instead of a string injects explicit noise (golden replays).
        ``record``: optional list; receives per-step dicts of device tensors (tests).
        """
        L = _lib.lib()
        if self.dry_run:
            raise _lib.CcdmError("dry-run engine cannot execute")
        if x.dim() == 4:
            B, _, H, W = x.shape
        else:
            B, H, W = x.shape
        noise_list = None
        if not isinstance(noise, str):
            noise_list = list(noise)
            if len(noise_list) != sum(1 for t in t_values if t > 1):
                raise ValueError("explicit noise needs one [B*H*W, K] tensor per step with t > 1")
            noise = "torch"
        elif noise not in ("torch", "philox"):
            raise ValueError(f"noise={noise!r}")
        n_lanes = min(self.lanes, B)
        if n_lanes > 1 and noise == "philox" and noise_list is None and record is None and img_rep == 1:
            return self._run_chain_lanes(n_lanes, x, condition, feature_condition, t_values, alphas, cumalphas, last_mode, seed, sample0)
        cur = _parent_stream if _parent_stream is not None else torch.cuda.current_stream(self.device)
        self.stream.wait_stream(cur)
        with torch.cuda.stream(self.stream):
            self.weights.refresh()
            prog = self.program(B, H, W, rows_per_sample=0, img_rep=img_rep)
            self._load_inputs(prog, x, condition, feature_condition)
            entries = []
            n = len(t_values)
            for i, t in enumerate(t_values):
                t0 = t - 1  # diffusion_denoising.py:107-113
                a, c = (0.0, 1.0) if t0 == 0 else (alphas[t0], cumalphas[t0 - 1])
                mode = _lib.DRAW_SAMPLE if t > 1 else last_mode  # :206-212
                entries.append((float(t), a, c, mode, i, i))
            self._write_tables(prog, entries, [float(t) for t in t_values])
            prog.run_pre(self._sp())
            use_tensor = noise == "torch"
            want_noise_out = record is not None and not use_tensor
            prog.set_noise(_lib.NOISE_TENSOR if use_tensor else _lib.NOISE_PHILOX, seed, sample0, use_noise_buf=use_tensor,
                           export_noise=want_noise_out)
            if not use_tensor and record is None:
                _lib.check(L.ccdm_plan_run(prog.plan, n, 1 if self.use_graph else 0, self._sp()), "plan_run")
            else:
                for i, t in enumerate(t_values):
                    if use_tensor and t > 1:
                        if noise_list is not None:
                            prog.noise.copy_(noise_list.pop(0).reshape(prog.noise.shape))
                        else:
                            prog.noise.exponential_(1)  # the reference's torch.multinomial draw
                    if record is not None:
                        rec = dict(t=t, labels_in=prog.labels.clone())
                    _lib.check(L.ccdm_plan_step(prog.plan, 1 if self.use_graph else 0, self._sp()), "plan_step")
                    if record is not None:
                        lg = prog.tens["logits"]
                        rec.update(labels_out=prog.labels.clone(), noise=prog.noise.clone() if t > 1 else None,
                                   logits=self.tensor_view(prog, lg).clone())
                        record.append(rec)
            labels = prog.labels.clone()
            probs = prog.probs.clone() if last_mode == _lib.DRAW_CONFIDENCE and t_values[-1] == 1 else None
        if _parent_stream is None:
            cur.wait_stream(self.stream)
        return labels, probs

    @torch.no_grad()
    def trace_step(self, x, condition, feature_condition, t: float, alpha=0.0, cumalpha=1.0, mode=_lib.DRAW_X0):
        """Debug/test aid: run ONE step op by op and return {tensor name: NHWC fp32 copy} of every
        op output (the liveness-planned workspace recycles buffers, so a finished step only holds
        the tail).  Same kernels, same order as the captured graph."""
        L = _lib.lib()
        if x.dim() == 4:
            B, _, H, W = x.shape
        else:
            B, H, W = x.shape
        out = {}
        cur = torch.cuda.current_stream(self.device)
        self.stream.wait_stream(cur)
        with torch.cuda.stream(self.stream):
            self.weights.refresh()
            prog = self.program(B, H, W, rows_per_sample=0)
            self._load_inputs(prog, x, condition, feature_condition)
            self._write_tables(prog, [(float(t), alpha, cumalpha, mode, 0, 0)], [float(t)])
            prog.run_pre(self._sp())
            prog.set_noise(_lib.NOISE_PHILOX)
            for i, o in enumerate(prog._op_dicts):
                _lib.check(L.ccdm_launch_op(ctypes.byref(prog._op_array[i]), self._sp()), f"op {i}")
                ten = o["_out"]
                if ten is not None:
                    out[ten.name] = self.tensor_view(prog, ten).float().clone()
                    if ten.want_stat and ten.stat_layout is None:
                        soff = ten.stat_addr - prog.workspace.data_ptr()
                        out[ten.name + "#stat"] = prog.workspace[soff:soff + B * ten.C * 16].view(torch.float64).view(B, ten.C, 2).clone()
            out["probs"] = prog.probs.clone()
            out["labels"] = prog.labels.clone()
        cur.wait_stream(self.stream)
        torch.cuda.synchronize(self.device)
        return out

    def tensor_view(self, prog: Program, t: Ten) -> torch.Tensor:
        """NHWC view (fp32 tensors) or NHWC copy (bf16 / fp16x2 tensors, stored plane-major) of a workspace tensor."""
        off = t.addr - prog.workspace.data_ptr()
        n = prog.B * t.H * t.W * t.C * t.esize
        raw = prog.workspace[off:off + n]
        if t.fmt == _lib.DT_F32:
            return raw.view(torch.float32).view(prog.B, t.H, t.W, t.C)
        if t.fmt == _lib.DT_F16X2:
            return from_pm_x3(raw.view(torch.float16).view(prog.B, t.C // 8, 2, t.H, t.W, 8))
        return from_pm(raw.view(torch.bfloat16).view(prog.B, t.C // 8, t.H, t.W, 8))
