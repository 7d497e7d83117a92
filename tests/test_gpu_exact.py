"""Kernel-level parity of the EXACT tensor-core mode (CCDM_DT_F16X2: every operand carried as fp16 hi + lo, three
tcgen05 MMAs per product, fp32 accumulation) on the B200, through the C ABI.

Yardstick: the same fused op in float64 on the CPU.  Stated tolerance: the kernel's error against float64 is at most
X3_REL = 5e-6 of the output's scale (max |ref|) -- fp32-grade: torch's own fp32 evaluation of the same op sits at
2e-7 (printed next to ours), a sequential fp32 sum of the same 288..4032 products at ~1e-6, the bf16 tensor-core mode
at 5e-3.  What is left is the tensor core's accumulator: it truncates on every accumulating MMA, so the error grows with
the number of K steps (measured 0.9e-6 at Cin = 96, 3.2e-6 at Cin = 448) -- not with the operand split, whose dropped
lo*lo term is 2^-22.
"""
import ctypes
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

X3_REL = 5e-6


@pytest.fixture(scope="module")
def L(cuda_device):
    from ccdm_b200 import _lib
    _lib.require_device()
    return _lib.lib()


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


def _check(out, args, kw, ostat=None, what=""):
    from gpu_util import ref_conv
    ref64 = ref_conv(*args, dtype=torch.float64, **kw)
    ref32 = ref_conv(*args, **kw)
    got = out.double().permute(0, 3, 1, 2).cpu()
    scale = float(ref64.abs().max())
    err = float((got - ref64).abs().max()) / scale
    err32 = float((ref32.double() - ref64).abs().max()) / scale
    print(f"x3 {what}: rel err {err:.2e} (torch fp32: {err32:.2e})")
    assert err <= X3_REL, (what, err, err32)
    if ostat is not None:
        s1 = ref64.sum(dim=(2, 3))
        s2 = (ref64 * ref64).sum(dim=(2, 3))
        n = ref64.shape[2] * ref64.shape[3]
        rms = (s2 / n).sqrt()
        # sums of fp32 values accumulated in fp32 per thread / warp / CTA and folded in double; the accumulator truncation of the
        # tensor core is one-sided, so the error of a sum is the error of its terms (X3_REL), not a random walk
        assert float(((ostat[..., 0].cpu() - s1).abs() / (n * rms + 1e-9)).max()) < X3_REL
        assert float(((ostat[..., 1].cpu() - s2).abs() / (n * rms * rms + 1e-9)).max()) < 1e-5


@pytest.mark.parametrize("B,C0,C1,Cout,H,W", [(2, 32, 0, 32, 128, 128), (1, 32, 32, 32, 64, 64), (2, 64, 32, 64, 32, 32),
                                               (2, 128, 96, 96, 16, 16), (3, 128, 128, 128, 8, 8), (1, 64, 384, 64, 32, 64),
                                               (1, 32, 0, 32, 20, 72), (1, 32, 32, 32, 40, 200)])
def test_x3_conv3x3_gn_silu_concat(L, B, C0, C1, Cout, H, W):
    from gpu_util import X3, nhwc, run_conv
    xs = [_rand(B, C0, H, W, seed=1) * 1.5 + 0.3] + ([_rand(B, C1, H, W, seed=2) * 0.7 - 0.2] if C1 else [])
    cin = C0 + C1
    w, b = _rand(Cout, cin, 3, 3, seed=3) / math.sqrt(9 * cin), _rand(Cout, seed=4) * 0.1
    gn = (1 + 0.1 * _rand(cin, seed=5), 0.1 * _rand(cin, seed=6))
    emb = _rand(B, Cout, seed=7)
    out, ostat = run_conv([nhwc(x) for x in xs], w, b, gn=gn, silu=True, emb=emb, dtype=X3, tc=True)
    _check(out, (xs, w, b), dict(gn=gn, silu=True, emb=emb), ostat, f"conv3x3 {cin}->{Cout} @{H}x{W}")


def test_x3_second_half_skip_identity_upsample(L):
    from gpu_util import X3, nhwc, run_conv
    B, C0, C1, Cout, H, W = 2, 64, 32, 64, 32, 32
    h1 = _rand(B, Cout, H, W, seed=21)
    xa, xb = _rand(B, C0, H, W, seed=22), _rand(B, C1, H, W, seed=23)
    w, b = _rand(Cout, Cout, 3, 3, seed=24) / math.sqrt(9 * Cout), _rand(Cout, seed=25) * 0.1
    ws, bs = _rand(Cout, C0 + C1, 1, 1, seed=26) / math.sqrt(C0 + C1), _rand(Cout, seed=27) * 0.1
    gn = (1 + 0.1 * _rand(Cout, seed=28), 0.1 * _rand(Cout, seed=29))
    out, ostat = run_conv([nhwc(h1)], w, b + bs, gn=gn, silu=True, skip=[nhwc(xa), nhwc(xb)], skip_w=ws, dtype=X3, tc=True)
    _check(out, ([h1], w, b), dict(gn=gn, silu=True, skip=[xa, xb], skip_w=ws, skip_b=bs), ostat, "res second half + 1x1 skip")
    # identity residual as an MMA chunk with the 2^shift identity matrix (what the engine binds in this mode)
    x = _rand(B, Cout, H, W, seed=30) * 3.0
    eye = torch.eye(Cout).reshape(Cout, Cout, 1, 1)
    out, ostat = run_conv([nhwc(h1)], w, b, gn=gn, silu=True, skip=[nhwc(x)], skip_w=eye, dtype=X3, tc=True)
    _check(out, ([h1], w, b), dict(gn=gn, silu=True, res=x), ostat, "res second half + identity chunk")
    out, ostat = run_conv([nhwc(h1)], w, b, upsample=True, dtype=X3, tc=True)
    _check(out, ([h1], w, b), dict(upsample=True), ostat, "upsample")


@pytest.mark.parametrize("B,C,Cs,H,W", [(3, 192, 128, 16, 16), (2, 256, 192, 8, 8), (1, 160, 96, 16, 16)])
def test_x3_wide_second_half_with_fused_skip(L, B, C, Cs, H, W):
    """More than 128 output channels (base_channels = 64 nets): the fused skip chunks follow the 3x3 conv's N tile."""
    from gpu_util import X3, nhwc, run_conv
    h1, xs = _rand(B, C, H, W, seed=21), _rand(B, Cs, H, W, seed=22)
    w, b = _rand(C, C, 3, 3, seed=24) / math.sqrt(9 * C), _rand(C, seed=25) * 0.1
    ws, bs = _rand(C, Cs, 1, 1, seed=26) / math.sqrt(Cs), _rand(C, seed=27) * 0.1
    gn = (1 + 0.1 * _rand(C, seed=28), 0.1 * _rand(C, seed=29))
    out, ostat = run_conv([nhwc(h1)], w, b + bs, gn=gn, silu=True, skip=[nhwc(xs)], skip_w=ws, dtype=X3, tc=True)
    _check(out, ([h1], w, b), dict(gn=gn, silu=True, skip=[xs], skip_w=ws, skip_b=bs), ostat, f"wide res {C} + 1x1 skip {Cs}")
    x = _rand(B, C, H, W, seed=30) * 3.0
    eye = torch.eye(C).reshape(C, C, 1, 1)
    out, ostat = run_conv([nhwc(h1)], w, b, gn=gn, silu=True, skip=[nhwc(x)], skip_w=eye, dtype=X3, tc=True)
    _check(out, ([h1], w, b), dict(gn=gn, silu=True, res=x), ostat, f"wide res {C} + identity chunk")
    wp, bp = _rand(C, C, 1, 1, seed=37) / math.sqrt(C), _rand(C, seed=38) * 0.1   # attention proj_out + x
    out, ostat = run_conv([nhwc(h1)], wp, bp, ksize=1, skip=[nhwc(x)], skip_w=eye, dtype=X3, tc=True)
    _check(out, ([h1], wp, bp), dict(res=x), ostat, f"wide proj {C} + residual")


@pytest.mark.parametrize("B,C,H,W", [(2, 32, 128, 128), (3, 64, 32, 32), (2, 96, 16, 16), (1, 128, 8, 16), (1, 32, 21, 37)])
def test_x3_downsample_stride2(L, B, C, H, W):
    from gpu_util import X3, nhwc, run_conv
    x = _rand(B, C, H, W, seed=41)
    w, b = _rand(C, C, 3, 3, seed=42) / math.sqrt(9 * C), _rand(C, seed=43) * 0.1
    out, ostat = run_conv([nhwc(x)], w, b, stride=2, dtype=X3, tc=True)
    assert out.shape == (B, (H + 1) // 2, (W + 1) // 2, C)
    _check(out, ([x], w, b), dict(stride=2), ostat, f"downsample {C} @{H}x{W}")


def test_x3_conv1x1_qkv_proj_and_head(L):
    from gpu_util import X3, nhwc, run_conv
    for (C, H, W) in [(96, 16, 16), (128, 8, 8), (64, 32, 64)]:
        B = 2
        x = _rand(B, C, H, W, seed=31)
        w, b = _rand(3 * C, C, 1, 1, seed=32) / math.sqrt(C), _rand(3 * C, seed=33) * 0.1
        gn = (1 + 0.1 * _rand(C, seed=34), 0.1 * _rand(C, seed=35))
        out, _ = run_conv([nhwc(x)], w, b, gn=gn, silu=False, ksize=1, want_stat=False, dtype=X3, tc=True)
        _check(out, ([x], w, b), dict(gn=gn), None, f"qkv {C}")
        a = _rand(B, C, H, W, seed=36)
        wp, bp = _rand(C, C, 1, 1, seed=37) / math.sqrt(C), _rand(C, seed=38) * 0.1
        eye = torch.eye(C).reshape(C, C, 1, 1)
        out, ostat = run_conv([nhwc(a)], wp, bp, ksize=1, skip=[nhwc(x)], skip_w=eye, dtype=X3, tc=True)
        _check(out, ([a], wp, bp), dict(res=x), ostat, f"proj + residual {C}")
    for K in (2, 20):  # output head: ragged Cout, fp32 logits
        x = _rand(2, 32, 64, 64, seed=51)
        w, b = _rand(K, 32, 3, 3, seed=52) / math.sqrt(9 * 32), _rand(K, seed=53) * 0.1
        gn = (1 + 0.1 * _rand(32, seed=54), 0.1 * _rand(32, seed=55))
        out, _ = run_conv([nhwc(x)], w, b, gn=gn, silu=True, want_stat=False, out_f32=True, dtype=X3, tc=True)
        assert out.dtype == torch.float32
        _check(out, ([x], w, b), dict(gn=gn, silu=True), None, f"head K={K}")


def test_x3_small_and_large_magnitudes(L):
    """The (hi, lo) split keeps 22 bits over the whole useful range: inputs of scale 1e-2 and 30, weights of scale 1e-3."""
    from gpu_util import X3, nhwc, run_conv
    B, C, H, W = 1, 64, 16, 16
    for xs, ws in ((1e-2, 1.0), (30.0, 1.0), (1.0, 1e-3)):
        x = _rand(B, C, H, W, seed=71) * xs
        w, b = _rand(C, C, 3, 3, seed=72) * ws / math.sqrt(9 * C), _rand(C, seed=73) * 0.1 * xs * ws
        out, _ = run_conv([nhwc(x)], w, b, stride=2, dtype=X3, tc=True, want_stat=False)  # raw operand path (no norm)
        _check(out, ([x], w, b), dict(stride=2), None, f"magnitudes x{xs} w{ws}")


def test_x3_encode_input_and_feature_planes(L):
    """one-hot(labels) ++ image and the NCHW feature condition as fp16x2 planes: hi + lo == 16 * value to 2^-22."""
    from ccdm_b200 import _lib
    from ccdm_b200.engine import from_pm_x3
    from gpu_util import sp
    for (B, K, C_img, H, W) in [(2, 2, 1, 16, 24), (1, 20, 3, 8, 8)]:
        CP = (K + C_img + 15) // 16 * 16
        g = torch.Generator().manual_seed(5)
        labels = torch.randint(0, K, (B, H, W), generator=g, dtype=torch.uint8).cuda()
        image = (torch.randn((B, C_img, H, W), generator=g) * 3).cuda()
        out = torch.full((B, CP // 8, 2, H, W, 8), float("nan"), dtype=torch.float16, device="cuda")
        op = _lib.Op(kind=_lib.OP_ENCODE_INPUT, dtype=_lib.DT_F16X2, out_dtype=_lib.DT_F16X2, B=B, Hin=H, Win=W, Hout=H, Wout=W,
                     Cout=CP, K=K, C_img=C_img)
        op.labels_in, op.image, op.out = labels.data_ptr(), image.data_ptr(), out.data_ptr()
        _lib.check(L.ccdm_launch_op(ctypes.byref(op), sp()))
        torch.cuda.synchronize()
        got = from_pm_x3(out)
        want = torch.zeros((B, H, W, CP), device="cuda")
        want[..., :K] = torch.nn.functional.one_hot(labels.long(), K).float()
        want[..., K:K + C_img] = image.permute(0, 2, 3, 1)
        assert float((got - want).abs().max()) <= 2.0 ** -22 * float(want.abs().max())
    B, C, H, W = 2, 384, 8, 16
    x = _rand(B, C, H, W, seed=9).cuda() * 2
    dst = torch.full((B, C // 8, 2, H, W, 8), float("nan"), dtype=torch.float16, device="cuda")
    stat = torch.zeros(B, C, 2, dtype=torch.float64, device="cuda")
    _lib.check(L.ccdm_nchw_to_nhwc_stats(x.data_ptr(), B, C, H, W, _lib.DT_F16X2, dst.data_ptr(), stat.data_ptr(), 1, sp()))
    torch.cuda.synchronize()
    assert float((from_pm_x3(dst) - x.permute(0, 2, 3, 1)).abs().max()) <= 2.0 ** -22 * float(x.abs().max())
    xd = x.double()
    assert torch.allclose(stat[..., 0], xd.sum(dim=(2, 3)), rtol=0, atol=1e-9)
    assert torch.allclose(stat[..., 1], (xd * xd).sum(dim=(2, 3)), rtol=0, atol=1e-9)


@pytest.mark.parametrize("B,heads,T", [(2, 3, 256), (3, 4, 64), (1, 2, 2048), (2, 4, 100), (1, 4, 512), (2, 1, 129)])
def test_x3_attention(L, B, heads, T):
    """QKVAttentionLegacy (unet.py:343-360) on the tensor cores with fp16x2 operands vs float64: <= 3e-6 of the output scale
    (the approximate exp2 of the online softmax is good to 2^-22; torch fp32 two-pass softmax: ~5e-7)."""
    from ccdm_b200 import _lib
    from ccdm_b200.engine import from_pm_x3, to_pm_x3
    from gpu_util import sp
    D = 32
    C = heads * D
    qkv = _rand(B, 3 * C, T, seed=63) * 1.3
    q, k, v = qkv.double().reshape(B * heads, 3 * D, T).split(D, dim=1)
    s = 1 / math.sqrt(math.sqrt(D))
    wgt = torch.softmax(torch.einsum("bct,bcs->bts", q * s, k * s), dim=-1)
    ref = torch.einsum("bts,bcs->bct", wgt, v).reshape(B, C, T)
    src = to_pm_x3(qkv.permute(0, 2, 1).reshape(B, 1, T, 3 * C).cuda())  # [B, 3C/8, 2, 1, T, 8]
    out = torch.full((B, C // 8, 2, 1, T, 8), float("nan"), dtype=torch.float16, device="cuda")
    op = _lib.Op(kind=_lib.OP_ATTENTION, dtype=_lib.DT_F16X2, out_dtype=_lib.DT_F16X2, B=B, Hin=1, Win=T, Hout=1, Wout=T,
                 C0=3 * C, Cout=C, heads=heads, head_dim=D, exact=0)
    op.src0, op.out = src.data_ptr(), out.data_ptr()
    _lib.check(L.ccdm_launch_op(ctypes.byref(op), sp()))
    torch.cuda.synchronize()
    got = from_pm_x3(out).reshape(B, T, C).double().cpu().permute(0, 2, 1)
    err = float((got - ref).abs().max()) / float(ref.abs().max())
    print(f"x3 attention T={T}: rel err {err:.2e}")
    assert err < 3e-6, err
