"""Helpers for the GPU parity tests: build single ops through the C ABI on torch-owned buffers."""
import ctypes

import numpy as np
import torch
import torch.nn.functional as F

from ccdm_b200 import _lib
from ccdm_b200.engine import (from_pm, from_pm_x3, pack_bias, pack_conv_weight, pack_conv_weight_tc, pack_conv_weight_x3,
                              subpixel_weights, to_pm, to_pm_x3)

X3 = "x3"  # run_conv(dtype=X3): fp16x2 storage (CCDM_DT_F16X2), the exact tensor-core mode


def x3_shift(*ws):
    """Power-of-two weight scale the engine would pick (PackedWeights.refresh) for these weight tensors."""
    import math
    wmax = max(float(w.abs().max()) for w in ws if w is not None)
    return int(max(0, min(13, math.floor(math.log2(32768.0 / max(wmax, 1e-30))))))


def sp():
    return _lib.stream_ptr(torch.cuda.current_stream())


def nhwc(x, dtype=torch.float32):
    """NCHW cpu/gpu -> contiguous NHWC on the GPU."""
    return x.permute(0, 2, 3, 1).contiguous().to("cuda", dtype)


def stats_of(x_nhwc):
    """double [B,C,2] = (sum, sum of squares) over pixels, as the producing kernel would emit."""
    xd = x_nhwc.double()
    return torch.stack([xd.sum(dim=(1, 2)), (xd * xd).sum(dim=(1, 2))], dim=-1).contiguous()


class StepCtx:
    """A one-row step table + counter on the device."""

    def __init__(self, t=0.0, alpha=0.0, cum=1.0, mode=_lib.DRAW_X0, draw=0, emb_row=0):
        e = _lib.StepEntry(float(t), float(alpha), float(cum), int(mode), int(draw), int(emb_row), 0, 0)
        self.table = torch.frombuffer(bytearray(bytes(e)), dtype=torch.uint8).cuda()
        self.counter = torch.zeros(1, dtype=torch.int32, device="cuda")


def run_conv(srcs, weight, bias, *, gn=None, silu=False, ksize=3, stride=1, upsample=False, skip=None, skip_w=None,
             res=None, emb=None, dtype=torch.float32, out_f32=False, want_stat=True, labels=None, image=None, K=0, tc=False):
    """Launch one CCDM_OP_CONV.  srcs: list of NHWC GPU tensors (1 or 2) or [] with labels/image.
    weight: OIHW fp32 (cpu or gpu); gn: (gamma, beta) or None; skip: list of NHWC tensors for the fused
    1x1 skip conv with skip_w [Cout, Cin_skip, 1, 1]; res: NHWC residual; emb: fp32 [B or 1, Cout] rows.
    Returns (out NHWC, ostat or None)."""
    L = _lib.lib()
    dev = "cuda"
    one_hot_in = labels is not None
    # bf16 activations live in the kernels' plane-major layout; this helper keeps NHWC at its surface
    x3 = isinstance(dtype, str) and dtype == X3
    pm = x3 or dtype == torch.bfloat16
    shapes = [tuple(s.shape) for s in srcs]
    skip_ch = [int(t.shape[3]) for t in skip] if skip is not None else []
    if x3:
        assert tc and res is None
        nh = list(srcs)
        srcs = [to_pm_x3(s.float()) for s in srcs]
        skip = [to_pm_x3(t.float()) for t in skip] if skip is not None else None
    elif pm:
        nh = list(srcs)
        srcs = [to_pm(s) for s in srcs]
        skip = [to_pm(s) for s in skip] if skip is not None else None
        res = to_pm(res) if res is not None else None
    if one_hot_in:
        B, Hin, Win = labels.shape
        C_img = image.shape[1]
        cin = K + C_img
    else:
        B, Hin, Win, _ = shapes[0]
        cin = sum(sh[3] for sh in shapes)
    Cout = weight.shape[0]
    Hout = Hin * 2 if upsample else ((Hin + 1) // 2 if stride == 2 else Hin)
    Wout = Win * 2 if upsample else ((Win + 1) // 2 if stride == 2 else Win)
    keep = []
    shift = 0
    if x3:
        wraw = subpixel_weights(weight.to(dev)) if upsample else weight.to(dev)
        shift = x3_shift(wraw, skip_w)
        wp = pack_conv_weight_x3(wraw, shift).contiguous()
    elif tc:
        wp = pack_conv_weight_tc(subpixel_weights(weight.to(dev)) if upsample else weight.to(dev)).contiguous()
    else:
        wp = pack_conv_weight(weight.to(dev), (cin + 7) // 8 * 8 if one_hot_in else None).contiguous()
    b = bias.to(dev).float()
    if skip is not None:
        assert skip_w is not None
    bp = pack_bias(b)
    out_dtype = torch.float32 if (out_f32 or (not x3 and dtype == torch.float32)) else (torch.float16 if x3 else torch.bfloat16)
    out = torch.full((B, Hout, Wout, Cout * (2 if out_dtype == torch.float16 else 1)), float("nan"), dtype=out_dtype, device=dev)
    dt_code = _lib.DT_F16X2 if x3 else (_lib.DT_F32 if dtype == torch.float32 else _lib.DT_BF16)
    op = _lib.Op(kind=_lib.OP_CONV, dtype=dt_code, acc_shift=shift + _lib.F16X2_SCALE_LOG2 if x3 else 0,
                 out_dtype=_lib.DT_F32 if out_dtype == torch.float32 else dt_code, B=B, Hin=Hin, Win=Win, Hout=Hout,
                 Wout=Wout, Cout=Cout, ksize=ksize, stride=stride, upsample=int(upsample), gn=int(gn is not None),
                 silu=int(silu), exact=0 if tc else 1)
    op.weight, op.bias, op.out = wp.data_ptr(), bp.data_ptr(), out.data_ptr()
    if one_hot_in:
        op.kind = _lib.OP_INPUT_CONV
        op.src_kind, op.K, op.C_img = 1, K, C_img
        lab = labels.to(dev, torch.uint8).contiguous()
        img = image.to(dev, torch.float32).contiguous()
        keep += [lab, img]
        op.labels_in, op.image = lab.data_ptr(), img.data_ptr()
    else:
        op.src0, op.C0 = srcs[0].data_ptr(), shapes[0][3]
        if len(srcs) > 1:
            op.src1, op.C1 = srcs[1].data_ptr(), shapes[1][3]
        if gn is not None:
            st = [stats_of(s) for s in (nh if pm else srcs)]
            keep += st
            op.stat0 = st[0].data_ptr()
            if len(srcs) > 1:
                op.stat1 = st[1].data_ptr()
            g, be = gn[0].to(dev).float().contiguous(), gn[1].to(dev).float().contiguous()
            keep += [g, be]
            op.gamma, op.beta = g.data_ptr(), be.data_ptr()
    if skip is not None:
        # the fused skip conv is an extra K chunk of THIS conv: its weights follow this conv's N tile
        nt = int(L.ccdm_conv_tc_nt(Cout, 16 if upsample else ksize * ksize, 1 if x3 else 0))
        sw = (pack_conv_weight_x3(skip_w.to(dev), shift, nt).contiguous() if x3 else
              pack_conv_weight_tc(skip_w.to(dev), nt).contiguous() if tc else skip_w.to(dev).float()[:, :, 0, 0].t().contiguous())
        keep.append(sw)
        op.skip0, op.S0 = skip[0].data_ptr(), skip_ch[0]
        if len(skip) > 1:
            op.skip1, op.S1 = skip[1].data_ptr(), skip_ch[1]
        op.skip_w = sw.data_ptr()
    if res is not None:
        op.res = res.data_ptr()
    ctx = StepCtx()
    keep.append(ctx)
    op.steps, op.step_ptr = ctx.table.data_ptr(), ctx.counter.data_ptr()
    if emb is not None:
        e = emb.to(dev).float().contiguous()
        keep.append(e)
        op.emb, op.emb_off, op.emb_cols = e.data_ptr(), 0, e.shape[1]
        op.emb_bstride = 1 if e.shape[0] == B and B > 1 else 0
    ostat = None
    if tc:
        assert L.ccdm_conv_uses_tc(ctypes.byref(op)) == 1, "op was expected to dispatch to the tcgen05 kernel"
    if want_stat:
        ostat = torch.full((B, Cout, 2), float("nan"), dtype=torch.float64, device=dev)
        op.ostat = 1
        part = torch.zeros(L.ccdm_op_part_floats(ctypes.byref(op)), dtype=torch.float32, device=dev)
        ticket = torch.zeros(B, dtype=torch.int32, device=dev)
        keep += [part, ticket]
        op.ostat, op.part, op.ticket = ostat.data_ptr(), part.data_ptr(), ticket.data_ptr()
    _lib.check(L.ccdm_launch_op(ctypes.byref(op), sp()), "conv")
    torch.cuda.synchronize()
    if want_stat:
        assert int(ticket.abs().sum()) == 0, "ticket counters must self-reset"
    if out_dtype == torch.float16:
        out = from_pm_x3(out.view(B, Cout // 8, 2, Hout, Wout, 8))
    elif pm and out_dtype == torch.bfloat16:
        out = from_pm(out.view(B, Cout // 8, Hout, Wout, 8))
    return out, ostat


def ref_conv(srcs_nchw, weight, bias, *, gn=None, silu=False, stride=1, upsample=False, skip=None, skip_w=None,
             skip_b=None, res=None, emb=None, dtype=torch.float32):
    """CPU torch reference of the same fused op (NCHW), fp32 by default (dtype=torch.float64: the yardstick both the
    kernel and the fp32 reference are measured against)."""
    cast = (lambda t: t.to(dtype).cpu()) if dtype != torch.float32 else None
    if cast is not None:
        f = lambda t: None if t is None else cast(t)  # noqa: E731
        return _ref_conv64([f(t) for t in srcs_nchw], f(weight), f(bias), gn=None if gn is None else (f(gn[0]), f(gn[1])), silu=silu,
                           stride=stride, upsample=upsample, skip=None if skip is None else [f(t) for t in skip], skip_w=f(skip_w),
                           skip_b=f(skip_b), res=f(res), emb=f(emb))
    x = torch.cat([s.float().cpu() for s in srcs_nchw], dim=1)
    h = x
    if gn is not None:
        h = F.group_norm(h, 32, gn[0].float().cpu(), gn[1].float().cpu(), eps=1e-5)
    if silu:
        h = F.silu(h)
    if upsample:
        h = F.interpolate(h, scale_factor=2, mode="nearest")
    h = F.conv2d(h, weight.float().cpu(), bias.float().cpu(), stride=stride, padding=weight.shape[-1] // 2)
    if emb is not None:
        h = h + emb.float().cpu()[:, :, None, None]
    if skip is not None:
        h = h + F.conv2d(torch.cat([s.float().cpu() for s in skip], dim=1), skip_w.float().cpu(), skip_b)
    if res is not None:
        h = h + res.float().cpu()
    return h


def _ref_conv64(srcs, weight, bias, *, gn, silu, stride, upsample, skip, skip_w, skip_b, res, emb):
    h = torch.cat(srcs, dim=1)
    if gn is not None:
        h = F.group_norm(h, 32, gn[0], gn[1], eps=1e-5)
    if silu:
        h = F.silu(h)
    if upsample:
        h = F.interpolate(h, scale_factor=2, mode="nearest")
    h = F.conv2d(h, weight, bias, stride=stride, padding=weight.shape[-1] // 2)
    if emb is not None:
        h = h + emb[:, :, None, None]
    if skip is not None:
        h = h + F.conv2d(torch.cat(skip, dim=1), skip_w, skip_b)
    if res is not None:
        h = h + res
    return h


def max_err(a, b):
    return float((a.double().cpu() - b.double().cpu()).abs().max())
