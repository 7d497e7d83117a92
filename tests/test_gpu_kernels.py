"""Kernel-level parity on the B200, through the C ABI (ccdm_launch_op & friends).

Integer / index results (labels, Philox words) are compared bit for bit with the C oracle;
floating-point kernels against a plain PyTorch fp32 CPU reference of the same op with the
tolerance written at each assert.
"""
import ctypes
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L(cuda_device):
    from ccdm_b200 import _lib
    _lib.require_device()
    return _lib.lib()


def _sp():
    from ccdm_b200 import _lib
    return _lib.stream_ptr(torch.cuda.current_stream())


# ---------------------------------------------------------------------------------------
# posterior + draw: bit-exact against the C oracle and the reference fixtures
# ---------------------------------------------------------------------------------------
def _posterior_draw(L, theta, labels, a, c, mode, noise=None, philox=None):
    from ccdm_b200 import _lib
    B = theta.shape[0]
    K = theta.shape[-1]
    n_pix = int(np.asarray(theta[0]).size) // K
    th = torch.as_tensor(theta).cuda().contiguous()
    lab = torch.as_tensor(labels).cuda().contiguous()
    out_l = torch.full(lab.shape, 255, dtype=torch.uint8, device="cuda")
    out_p = torch.full(th.shape, float("nan"), device="cuda")
    out_e = torch.full(th.shape, float("nan"), device="cuda")
    nz = torch.as_tensor(noise).cuda().contiguous() if noise is not None else None
    seed, draw, s0 = philox if philox is not None else (0, 0, 0)
    _lib.check(L.ccdm_posterior_draw(th.data_ptr(), lab.data_ptr(), n_pix, B, K, float(a), float(c), mode,
                                     _lib.NOISE_TENSOR if noise is not None else _lib.NOISE_PHILOX,
                                     nz.data_ptr() if nz is not None else None, seed, draw, s0, out_l.data_ptr(),
                                     out_p.data_ptr(), out_e.data_ptr(), _sp()))
    torch.cuda.synchronize()
    return out_l.cpu().numpy(), out_p.cpu().numpy(), out_e.cpu().numpy()


@pytest.mark.parametrize("K", [2, 20])
@pytest.mark.parametrize("T", [250, 1000])
def test_posterior_bit_exact_vs_oracle_and_reference(L, K, T):
    from ccdm_b200 import _lib
    from oracle import cdm
    g, sch = golden("posterior.npz"), golden("schedules.npz")
    tag = f"K{K}_T{T}"
    theta, labels, ts, post = (g[tag + s] for s in ("_theta", "_labels", "_t", "_post"))
    for i, t in enumerate(ts):
        a, c = cdm.step_scalars(sch[f"cosine{T}_alphas"], sch[f"cosine{T}_cumalphas"], int(t))
        _, raw, _ = _posterior_draw(L, theta[i:i + 1], labels[i:i + 1], a, c, _lib.DRAW_POSTERIOR)
        oracle = cdm.posterior_closed(labels[i], theta[i], a, c)
        np.testing.assert_array_equal(raw[0], oracle)                        # bit-exact vs the oracle
        np.testing.assert_allclose(raw[0], post[i], rtol=0, atol=2e-6)       # vs the reference's O(K^2) einsum


@pytest.mark.parametrize("K", [2, 20])
def test_draw_bit_exact_vs_reference_fixture(L, K):
    """alpha=0, cumalpha=1 makes the posterior the identity, isolating clamp + normalise + race."""
    from ccdm_b200 import _lib
    from oracle import cdm
    g = golden("draw.npz")
    p = g[f"K{K}_probs_in"]
    lab0 = np.zeros(p.shape[:-1], np.uint8)
    noise = g[f"K{K}_noise"].reshape(p.shape)
    # the identity posterior is only exact for theta/1*1: check through the oracle on the same input
    post = np.stack([cdm.posterior_closed(lab0[b], p[b], 0.0, 1.0) for b in range(p.shape[0])])
    o_lab, o_pn = cdm.draw(post, noise, 0)
    l, pn, e = _posterior_draw(L, p, lab0, 0.0, 1.0, _lib.DRAW_SAMPLE, noise=noise)
    np.testing.assert_array_equal(l, o_lab)
    np.testing.assert_array_equal(pn, o_pn)
    np.testing.assert_array_equal(e, noise)
    np.testing.assert_array_equal(l, g[f"K{K}_sample_labels"])            # == reference sample()
    l, pn, _ = _posterior_draw(L, p, lab0, 0.0, 1.0, _lib.DRAW_MAJORITY)
    np.testing.assert_array_equal(l, g[f"K{K}_majority_labels"])          # == reference max_prob_sample()
    l, pn, _ = _posterior_draw(L, p, lab0, 0.0, 1.0, _lib.DRAW_CONFIDENCE)
    np.testing.assert_allclose(pn, g[f"K{K}_confidence"], rtol=3e-7, atol=0)  # == reference prob_sample()


@pytest.mark.parametrize("K", [2, 3, 7, 20, 32])
def test_full_step_posterior_draw_random(L, K):
    from ccdm_b200 import _lib
    from oracle import cdm
    rng = np.random.default_rng(K)
    B, n = 3, 1000  # ragged: not a multiple of the CTA size
    theta = rng.dirichlet(np.ones(K) * 0.3, size=(B, n)).astype(np.float32)
    labels = rng.integers(0, K, size=(B, n)).astype(np.uint8)
    noise = rng.exponential(size=(B, n, K)).astype(np.float32)
    a, c = np.float32(0.9731), np.float32(0.4127)
    post = cdm.posterior_closed(labels, theta, a, c).reshape(theta.shape)
    o_lab, o_pn = cdm.draw(post, noise, 0)
    l, pn, _ = _posterior_draw(L, theta, labels, a, c, _lib.DRAW_SAMPLE, noise=noise)
    np.testing.assert_array_equal(l, o_lab)
    np.testing.assert_array_equal(pn, o_pn)


def test_philox_words_and_uniform_labels_bit_exact(L):
    from ccdm_b200 import _lib
    from oracle import cdm
    seed, draw, s0, ns, npx = 0x0123456789ABCDEF, 17, 5, 3, 777
    for K in (2, 20):
        bits = torch.zeros((ns, npx, K), dtype=torch.int32, device="cuda")
        _lib.check(L.ccdm_philox_bits(seed, draw, s0, ns, npx, K, bits.data_ptr(), _sp()))
        ref = cdm.philox_bits(seed, draw, s0, ns, npx, K)
        np.testing.assert_array_equal(bits.cpu().numpy().view(np.uint32), ref)
        # the E the head kernel derives from those words, exported through noise_out
        theta = np.full((ns, npx, K), 1.0 / K, np.float32)
        lab = np.zeros((ns, npx), np.uint8)
        l, pn, e = _posterior_draw(L, theta, lab, 0.0, 1.0, _lib.DRAW_SAMPLE, philox=(seed, draw, s0))
        e_ref = cdm.bits_to_exponential(ref)
        np.testing.assert_allclose(e, e_ref, rtol=2e-7, atol=1e-7)   # logf: CUDA vs libm, <= 1-2 ulp
        o_lab, _ = cdm.draw(np.stack([cdm.posterior_closed(lab[b], theta[b], 0.0, 1.0) for b in range(ns)]), e, 0)
        np.testing.assert_array_equal(l, o_lab)                      # bit-exact given the exported E
        xt = torch.zeros((ns, npx), dtype=torch.uint8, device="cuda")
        _lib.check(L.ccdm_uniform_labels(seed, draw, s0, ns, npx, K, xt.data_ptr(), _sp()))
        agree = (xt.cpu().numpy() == cdm.uniform_labels(e_ref)).mean()
        assert agree > 0.9999
        counts = np.bincount(xt.cpu().numpy().ravel(), minlength=K) / (ns * npx)
        assert np.abs(counts - 1.0 / K).max() < 0.05


def test_sharding_invariance_of_philox(L):
    """Samples [4,8) drawn as a shard equal samples 4..7 of the full batch."""
    from ccdm_b200 import _lib
    K, npx = 20, 513
    full = torch.zeros((8, npx), dtype=torch.uint8, device="cuda")
    part = torch.zeros((4, npx), dtype=torch.uint8, device="cuda")
    _lib.check(L.ccdm_uniform_labels(99, 0, 0, 8, npx, K, full.data_ptr(), _sp()))
    _lib.check(L.ccdm_uniform_labels(99, 0, 4, 4, npx, K, part.data_ptr(), _sp()))
    assert torch.equal(full[4:], part)


def test_onehot_label_roundtrip(L):
    from ccdm_b200 import _lib
    B, K, H, W = 2, 20, 9, 13
    lab = torch.randint(0, K, (B, H, W), dtype=torch.uint8, device="cuda")
    oh = torch.zeros((B, H, W, K), dtype=torch.int64, device="cuda")
    _lib.check(L.ccdm_labels_to_onehot_i64(lab.data_ptr(), B * H * W, K, oh.data_ptr(), _sp()))
    assert torch.equal(oh, F.one_hot(lab.long(), K))
    x = oh.permute(0, 3, 1, 2).float()  # NHWC-strided BCHW view, like the reference's sample()
    back = torch.zeros_like(lab)
    _lib.check(L.ccdm_onehot_to_labels(x.data_ptr(), *x.stride(), B, K, H, W, back.data_ptr(), _sp()))
    assert torch.equal(back, lab)
    xc = x.contiguous()  # and plain NCHW
    _lib.check(L.ccdm_onehot_to_labels(xc.data_ptr(), *xc.stride(), B, K, H, W, back.data_ptr(), _sp()))
    assert torch.equal(back, lab)


# ---------------------------------------------------------------------------------------
# fused conv variants (fp32 exact kernels) vs torch fp32 CPU
# ---------------------------------------------------------------------------------------
CONV_TOL = 2e-4  # abs, outputs O(1): fp32 accumulation in a different order over <= 9*448 products


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


@pytest.mark.parametrize("B,C0,C1,Cout,H,W", [(2, 32, 0, 32, 16, 32), (1, 64, 32, 64, 24, 20), (2, 64, 384, 64, 8, 16),
                                               (1, 128, 96, 96, 8, 8), (3, 32, 0, 96, 5, 7)])
def test_conv3x3_gn_silu_concat(L, B, C0, C1, Cout, H, W):
    from gpu_util import max_err, nhwc, ref_conv, run_conv
    xs = [_rand(B, C0, H, W, seed=1) * 1.5 + 0.3] + ([_rand(B, C1, H, W, seed=2) * 0.7 - 0.2] if C1 else [])
    cin = C0 + C1
    w, b = _rand(Cout, cin, 3, 3, seed=3) / math.sqrt(9 * cin), _rand(Cout, seed=4) * 0.1
    gn = (1 + 0.1 * _rand(cin, seed=5), 0.1 * _rand(cin, seed=6))
    emb = _rand(B, Cout, seed=7)
    out, ostat = run_conv([nhwc(x) for x in xs], w, b, gn=gn, silu=True, emb=emb)
    ref = ref_conv(xs, w, b, gn=gn, silu=True, emb=emb)
    assert max_err(out.permute(0, 3, 1, 2), ref) < CONV_TOL
    rs = torch.stack([ref.double().sum(dim=(2, 3)), (ref.double() ** 2).sum(dim=(2, 3))], dim=-1)
    np.testing.assert_allclose(ostat.cpu().numpy(), rs.numpy(), rtol=2e-5, atol=2e-3)


@pytest.mark.parametrize("H,W", [(16, 32), (9, 11), (8, 8)])
def test_conv_downsample_and_upsample(L, H, W):
    from gpu_util import max_err, nhwc, ref_conv, run_conv
    B, C = 2, 64
    x = _rand(B, C, H, W, seed=11)
    w, b = _rand(C, C, 3, 3, seed=12) / math.sqrt(9 * C), _rand(C, seed=13) * 0.1
    out, _ = run_conv([nhwc(x)], w, b, stride=2)
    assert max_err(out.permute(0, 3, 1, 2), ref_conv([x], w, b, stride=2)) < CONV_TOL
    out, _ = run_conv([nhwc(x)], w, b, upsample=True)
    assert max_err(out.permute(0, 3, 1, 2), ref_conv([x], w, b, upsample=True)) < CONV_TOL


def test_conv_resblock_second_half_with_skip_and_residual(L):
    from gpu_util import max_err, nhwc, ref_conv, run_conv
    B, C0, C1, Cout, H, W = 2, 64, 32, 64, 16, 16
    h1 = _rand(B, Cout, H, W, seed=21)
    xa, xb = _rand(B, C0, H, W, seed=22), _rand(B, C1, H, W, seed=23)
    w, b = _rand(Cout, Cout, 3, 3, seed=24) / math.sqrt(9 * Cout), _rand(Cout, seed=25) * 0.1
    ws, bs = _rand(Cout, C0 + C1, 1, 1, seed=26) / math.sqrt(C0 + C1), _rand(Cout, seed=27) * 0.1
    gn = (1 + 0.1 * _rand(Cout, seed=28), 0.1 * _rand(Cout, seed=29))
    out, _ = run_conv([nhwc(h1)], w, b + bs, gn=gn, silu=True, skip=[nhwc(xa), nhwc(xb)], skip_w=ws)
    ref = ref_conv([h1], w, b, gn=gn, silu=True, skip=[xa, xb], skip_w=ws, skip_b=bs)
    assert max_err(out.permute(0, 3, 1, 2), ref) < CONV_TOL
    x = _rand(B, Cout, H, W, seed=30)
    out, _ = run_conv([nhwc(h1)], w, b, gn=gn, silu=True, res=nhwc(x))
    assert max_err(out.permute(0, 3, 1, 2), ref_conv([h1], w, b, gn=gn, silu=True, res=x)) < CONV_TOL


def test_conv1x1_gn_nosilu_and_proj_residual(L):
    from gpu_util import max_err, nhwc, ref_conv, run_conv
    B, C, H, W = 2, 96, 16, 16
    x = _rand(B, C, H, W, seed=31)
    w, b = _rand(3 * C, C, 1, 1, seed=32) / math.sqrt(C), _rand(3 * C, seed=33) * 0.1
    gn = (1 + 0.1 * _rand(C, seed=34), 0.1 * _rand(C, seed=35))
    out, _ = run_conv([nhwc(x)], w, b, gn=gn, silu=False, ksize=1, want_stat=False)
    assert max_err(out.permute(0, 3, 1, 2), ref_conv([x], w, b, gn=gn)) < CONV_TOL
    a = _rand(B, C, H, W, seed=36)
    wp, bp = _rand(C, C, 1, 1, seed=37) / math.sqrt(C), _rand(C, seed=38) * 0.1
    out, ostat = run_conv([nhwc(a)], wp, bp, ksize=1, res=nhwc(x))
    assert max_err(out.permute(0, 3, 1, 2), ref_conv([a], wp, bp, res=x)) < CONV_TOL


@pytest.mark.parametrize("K,C_img,H,W", [(2, 1, 32, 32), (20, 3, 16, 24)])
def test_input_conv_on_labels_and_image(L, K, C_img, H, W):
    from gpu_util import max_err, ref_conv, run_conv
    B, Cout = 2, 32
    g = torch.Generator().manual_seed(41)
    labels = torch.randint(0, K, (B, H, W), generator=g, dtype=torch.uint8)
    image = _rand(B, C_img, H, W, seed=42)
    w, b = _rand(Cout, K + C_img, 3, 3, seed=43) / math.sqrt(9 * (K + C_img)), _rand(Cout, seed=44) * 0.1
    out, ostat = run_conv([], w, b, labels=labels, image=image, K=K)
    x = torch.cat([F.one_hot(labels.long(), K).permute(0, 3, 1, 2).float(), image], dim=1)  # unet.py:760
    assert max_err(out.permute(0, 3, 1, 2), ref_conv([x], w, b)) < CONV_TOL


@pytest.mark.parametrize("K", [2, 20])
def test_head_conv_ragged_cout_fp32_logits(L, K):
    from gpu_util import max_err, nhwc, ref_conv, run_conv
    B, C, H, W = 2, 32, 16, 16
    x = _rand(B, C, H, W, seed=51)
    w, b = _rand(K, C, 3, 3, seed=52) / math.sqrt(9 * C), _rand(K, seed=53) * 0.1
    gn = (1 + 0.1 * _rand(C, seed=54), 0.1 * _rand(C, seed=55))
    out, _ = run_conv([nhwc(x)], w, b, gn=gn, silu=True, want_stat=False, out_f32=True)
    assert max_err(out.permute(0, 3, 1, 2), ref_conv([x], w, b, gn=gn, silu=True)) < CONV_TOL


# ---------------------------------------------------------------------------------------
# tcgen05 conv kernel (bf16 storage, fp32 accumulation in TMEM) vs torch fp32 on the same
# bf16-rounded inputs and weights.  Tolerance: the normalised+activated input is re-rounded
# to bf16 before the MMA (rel 2^-9) and SiLU uses tanh.approx: |err| <= 2.5e-2 on O(1)
# outputs, mean |err| <= 3e-3.
# ---------------------------------------------------------------------------------------
TC_MAX, TC_MEAN = 2.5e-2, 3e-3


def _bf(x):
    return x.to(torch.bfloat16).float()


def _tc_check(out, ref, ostat=None):
    got = out.float().permute(0, 3, 1, 2).cpu()
    err = (got - ref).abs()
    assert float(err.max()) < TC_MAX and float(err.mean()) < TC_MEAN, (float(err.max()), float(err.mean()))
    if ostat is not None:
        # The statistics are accumulated from the fp32 accumulators BEFORE the bf16 rounding of the store
        # (conv_tc_common.cuh).  Rounding errors are zero-mean with |e| <= 2^-9 |v|, so against the sums of
        # the stored values the difference is a random walk: ~ 2^-9 * rms * sqrt(N) (x 2 rms for squares).
        gd = out.double()
        n = gd.shape[1] * gd.shape[2]
        s1, s2 = gd.sum(dim=(1, 2)), (gd * gd).sum(dim=(1, 2))
        rms = (s2 / n).sqrt()
        tol1 = 6 * 2.0 ** -9 * rms * n ** 0.5 + 1e-3
        tol2 = 20 * 2.0 ** -9 * rms * rms * n ** 0.5 + 1e-3  # squares: heavier tails (sum of 2 v e)
        got1, got2 = ostat[..., 0], ostat[..., 1]
        assert bool(((got1 - s1).abs() <= tol1).all()), float(((got1 - s1).abs() / tol1).max())
        assert bool(((got2 - s2).abs() <= tol2).all()), float(((got2 - s2).abs() / tol2).max())


@pytest.mark.parametrize("B,C0,C1,Cout,H,W", [(2, 32, 0, 32, 128, 128), (1, 32, 32, 32, 64, 64), (2, 64, 32, 64, 32, 32),
                                               (2, 128, 96, 96, 16, 16), (3, 128, 128, 128, 8, 8), (1, 64, 384, 64, 32, 64),
                                               (1, 32, 0, 32, 20, 72)])
def test_tc_conv3x3_gn_silu_concat(L, B, C0, C1, Cout, H, W):
    from gpu_util import nhwc, ref_conv, run_conv
    xs = [_bf(_rand(B, C0, H, W, seed=1) * 1.5 + 0.3)] + ([_bf(_rand(B, C1, H, W, seed=2) * 0.7 - 0.2)] if C1 else [])
    cin = C0 + C1
    w, b = _bf(_rand(Cout, cin, 3, 3, seed=3) / math.sqrt(9 * cin)), _rand(Cout, seed=4) * 0.1
    gn = (1 + 0.1 * _rand(cin, seed=5), 0.1 * _rand(cin, seed=6))
    emb = _rand(B, Cout, seed=7)
    out, ostat = run_conv([nhwc(x, torch.bfloat16) for x in xs], w, b, gn=gn, silu=True, emb=emb, dtype=torch.bfloat16, tc=True)
    _tc_check(out, ref_conv(xs, w, b, gn=gn, silu=True, emb=emb), ostat)


def test_tc_conv_second_half_with_skip_residual_and_upsample(L):
    from gpu_util import nhwc, ref_conv, run_conv
    B, C0, C1, Cout, H, W = 2, 64, 32, 64, 32, 32
    h1 = _bf(_rand(B, Cout, H, W, seed=21))
    xa, xb = _bf(_rand(B, C0, H, W, seed=22)), _bf(_rand(B, C1, H, W, seed=23))
    w, b = _bf(_rand(Cout, Cout, 3, 3, seed=24) / math.sqrt(9 * Cout)), _rand(Cout, seed=25) * 0.1
    ws, bs = _bf(_rand(Cout, C0 + C1, 1, 1, seed=26) / math.sqrt(C0 + C1)), _rand(Cout, seed=27) * 0.1
    gn = (1 + 0.1 * _rand(Cout, seed=28), 0.1 * _rand(Cout, seed=29))
    bf = torch.bfloat16
    out, ostat = run_conv([nhwc(h1, bf)], w, b + bs, gn=gn, silu=True, skip=[nhwc(xa, bf), nhwc(xb, bf)], skip_w=ws, dtype=bf, tc=True)
    _tc_check(out, ref_conv([h1], w, b, gn=gn, silu=True, skip=[xa, xb], skip_w=ws, skip_b=bs), ostat)
    x = _bf(_rand(B, Cout, H, W, seed=30))
    out, ostat = run_conv([nhwc(h1, bf)], w, b, gn=gn, silu=True, res=nhwc(x, bf), dtype=bf, tc=True)
    _tc_check(out, ref_conv([h1], w, b, gn=gn, silu=True, res=x), ostat)
    out, ostat = run_conv([nhwc(h1, bf)], w, b, upsample=True, dtype=bf, tc=True)
    _tc_check(out, ref_conv([h1], w, b, upsample=True), ostat)


@pytest.mark.parametrize("B,C,Cs,H,W", [(3, 192, 128, 16, 16), (2, 256, 192, 8, 8), (1, 160, 96, 16, 16)])
def test_tc_wide_second_half_with_fused_skip(L, B, C, Cs, H, W):
    """More than 128 output channels (base_channels = 64 nets): three to five N tiles per sample, and the fused 1x1 skip
    conv / identity residual chunk follows the 3x3 conv's tile, not the (wider) tile of a stand-alone 1x1 conv."""
    from gpu_util import nhwc, ref_conv, run_conv
    bf = torch.bfloat16
    h1, xs = _bf(_rand(B, C, H, W, seed=21)), _bf(_rand(B, Cs, H, W, seed=22))
    w, b = _bf(_rand(C, C, 3, 3, seed=24) / math.sqrt(9 * C)), _rand(C, seed=25) * 0.1
    ws, bs = _bf(_rand(C, Cs, 1, 1, seed=26) / math.sqrt(Cs)), _rand(C, seed=27) * 0.1
    gn = (1 + 0.1 * _rand(C, seed=28), 0.1 * _rand(C, seed=29))
    out, ostat = run_conv([nhwc(h1, bf)], w, b + bs, gn=gn, silu=True, skip=[nhwc(xs, bf)], skip_w=ws, dtype=bf, tc=True)
    _tc_check(out, ref_conv([h1], w, b, gn=gn, silu=True, skip=[xs], skip_w=ws, skip_b=bs), ostat)
    x = _bf(_rand(B, C, H, W, seed=30))
    eye = torch.eye(C).reshape(C, C, 1, 1)
    out, ostat = run_conv([nhwc(h1, bf)], w, b, gn=gn, silu=True, skip=[nhwc(x, bf)], skip_w=eye, dtype=bf, tc=True)
    _tc_check(out, ref_conv([h1], w, b, gn=gn, silu=True, res=x), ostat)
    wp, bp = _bf(_rand(C, C, 1, 1, seed=37) / math.sqrt(C)), _rand(C, seed=38) * 0.1   # attention proj_out + x
    out, ostat = run_conv([nhwc(h1, bf)], wp, bp, ksize=1, skip=[nhwc(x, bf)], skip_w=eye, dtype=bf, tc=True)
    _tc_check(out, ref_conv([h1], wp, bp, res=x), ostat)


@pytest.mark.parametrize("B,C,H,W", [(2, 32, 128, 128), (3, 64, 32, 32), (2, 96, 16, 16), (1, 128, 8, 16), (1, 32, 21, 37)])
def test_tc_conv_downsample_stride2(L, B, C, H, W):
    """Downsample (unet.py:136-139) on the TMA-fed tensor-core kernel: four parity sub-images gathered by
    TMA boxes with element stride 2; odd sizes exercise the zero fill on the right / bottom."""
    from gpu_util import nhwc, ref_conv, run_conv
    from ccdm_b200 import _lib
    bf = torch.bfloat16
    x = _bf(_rand(B, C, H, W, seed=41))
    w, b = _bf(_rand(C, C, 3, 3, seed=42) / math.sqrt(9 * C)), _rand(C, seed=43) * 0.1
    out, ostat = run_conv([nhwc(x, bf)], w, b, stride=2, dtype=bf, tc=True)
    assert out.shape == (B, (H + 1) // 2, (W + 1) // 2, C)
    _tc_check(out, ref_conv([x], w, b, stride=2), ostat)


def test_encode_input_planes(L):
    """bf16 mode: one-hot(labels) ++ image as a zero-padded plane-major tensor (unet.py:760)."""
    from ccdm_b200 import _lib
    from ccdm_b200.engine import from_pm
    for (B, K, C_img, H, W) in [(2, 2, 1, 16, 24), (1, 20, 3, 8, 8)]:
        CP = (K + C_img + 15) // 16 * 16
        g = torch.Generator().manual_seed(5)
        labels = torch.randint(0, K, (B, H, W), generator=g, dtype=torch.uint8).cuda()
        image = torch.randn((B, C_img, H, W), generator=g).cuda()
        out = torch.full((B, CP // 8, H, W, 8), float("nan"), dtype=torch.bfloat16, device="cuda")
        op = _lib.Op(kind=_lib.OP_ENCODE_INPUT, dtype=_lib.DT_BF16, out_dtype=_lib.DT_BF16, B=B, Hin=H, Win=W, Hout=H, Wout=W, Cout=CP,
                     K=K, C_img=C_img)
        op.labels_in, op.image, op.out = labels.data_ptr(), image.data_ptr(), out.data_ptr()
        _lib.check(L.ccdm_launch_op(ctypes.byref(op), _sp()))
        torch.cuda.synchronize()
        got = from_pm(out).float()  # [B,H,W,CP]
        want = torch.zeros((B, H, W, CP), device="cuda")
        want[..., :K] = F.one_hot(labels.long(), K).float()
        want[..., K:K + C_img] = image.permute(0, 2, 3, 1).to(torch.bfloat16).float()
        assert torch.equal(got, want)


def test_tc_conv1x1_qkv_proj_and_head(L):
    from gpu_util import nhwc, ref_conv, run_conv
    bf = torch.bfloat16
    for (C, H, W) in [(96, 16, 16), (128, 8, 8), (64, 32, 64)]:
        B = 2
        x = _bf(_rand(B, C, H, W, seed=31))
        w, b = _bf(_rand(3 * C, C, 1, 1, seed=32) / math.sqrt(C)), _rand(3 * C, seed=33) * 0.1
        gn = (1 + 0.1 * _rand(C, seed=34), 0.1 * _rand(C, seed=35))
        out, _ = run_conv([nhwc(x, bf)], w, b, gn=gn, silu=False, ksize=1, want_stat=False, dtype=bf, tc=True)
        _tc_check(out, ref_conv([x], w, b, gn=gn))
        a = _bf(_rand(B, C, H, W, seed=36))
        wp, bp = _bf(_rand(C, C, 1, 1, seed=37) / math.sqrt(C)), _rand(C, seed=38) * 0.1
        out, ostat = run_conv([nhwc(a, bf)], wp, bp, ksize=1, res=nhwc(x, bf), dtype=bf, tc=True)
        _tc_check(out, ref_conv([a], wp, bp, res=x), ostat)
    for K in (2, 20):  # output head: ragged Cout, fp32 logits
        x = _bf(_rand(2, 32, 64, 64, seed=51))
        w, b = _bf(_rand(K, 32, 3, 3, seed=52) / math.sqrt(9 * 32)), _rand(K, seed=53) * 0.1
        gn = (1 + 0.1 * _rand(32, seed=54), 0.1 * _rand(32, seed=55))
        out, _ = run_conv([nhwc(x, bf)], w, b, gn=gn, silu=True, want_stat=False, out_f32=True, dtype=bf, tc=True)
        assert out.dtype == torch.float32
        _tc_check(out, ref_conv([x], w, b, gn=gn, silu=True))


@pytest.mark.parametrize("K", [2, 20])
def test_head_fast_sampling_agrees_with_exact(L, K):
    """OP_HEAD with exact=0 (bf16 engine mode: approximate exp2/log2/reciprocal, no final normalisation) draws the same
    labels as the exact path from the same logits and Philox bits, up to near-ties of the race scores."""
    import ctypes
    from ccdm_b200 import _lib
    from gpu_util import StepCtx, sp
    B, H, W = 3, 64, 96
    g = torch.Generator().manual_seed(77)
    logits = (torch.randn(B, H, W, K, generator=g) * 3).cuda().contiguous()
    lab_in = torch.randint(0, K, (B, H, W), generator=g, dtype=torch.uint8).cuda()
    outs = []
    for exact in (1, 0):
        ctx = StepCtx(t=17.0, alpha=0.97, cum=0.61, mode=_lib.DRAW_SAMPLE, draw=5)
        lab_out = torch.full((B, H, W), 255, dtype=torch.uint8, device="cuda")
        ticket = torch.zeros(4, dtype=torch.int32, device="cuda")
        op = _lib.Op(kind=_lib.OP_HEAD, dtype=_lib.DT_F32, out_dtype=_lib.DT_F32, B=B, Hin=H, Win=W, Hout=H, Wout=W, K=K, exact=exact,
                     noise_mode=_lib.NOISE_PHILOX, seed=1234, sample0=7)
        op.src0, op.labels_in, op.labels_out = logits.data_ptr(), lab_in.data_ptr(), lab_out.data_ptr()
        op.steps, op.step_ptr, op.ticket = ctx.table.data_ptr(), ctx.counter.data_ptr(), ticket.data_ptr()
        _lib.check(L.ccdm_launch_op(ctypes.byref(op), sp()))
        torch.cuda.synchronize()
        assert int(ctx.counter.item()) == 1  # the head advances the device step counter on both paths
        assert int(lab_out.max()) < K
        outs.append(lab_out.cpu())
    mismatch = float((outs[0] != outs[1]).float().mean())
    print("head fast/exact label mismatch K=%d: %.2e" % (K, mismatch))
    assert mismatch < 1e-4, mismatch
    # and the draw is a real draw: not the argmax of the posterior everywhere
    assert float((outs[1] != logits.argmax(-1).cpu()).float().mean()) > 0.01


# ---------------------------------------------------------------------------------------
# attention (QKVAttentionLegacy) vs torch
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,heads,T", [(2, 3, 256), (1, 4, 64), (1, 2, 2048), (2, 4, 100)])
def test_attention_legacy_order(L, B, heads, T):
    from ccdm_b200 import _lib
    D = 32
    C = heads * D
    qkv = _rand(B, 3 * C, T, seed=61) * 1.3
    # reference: unet.py:343-360
    q, k, v = qkv.reshape(B * heads, 3 * D, T).split(D, dim=1)
    s = 1 / math.sqrt(math.sqrt(D))
    wgt = torch.softmax(torch.einsum("bct,bcs->bts", q * s, k * s), dim=-1)
    ref = torch.einsum("bts,bcs->bct", wgt, v).reshape(B, C, T)
    src = qkv.permute(0, 2, 1).contiguous().cuda()  # [B, T, 3C]
    out = torch.full((B, T, C), float("nan"), device="cuda")
    op = _lib.Op(kind=_lib.OP_ATTENTION, dtype=_lib.DT_F32, out_dtype=_lib.DT_F32, B=B, Hin=1, Win=T, Hout=1, Wout=T,
                 C0=3 * C, Cout=C, heads=heads, head_dim=D, exact=1)
    op.src0, op.out = src.data_ptr(), out.data_ptr()
    _lib.check(L.ccdm_launch_op(ctypes.byref(op), _sp()))
    torch.cuda.synchronize()
    err = float((out.cpu().permute(0, 2, 1) - ref).abs().max())
    assert err < 2e-5, err  # fp32, online softmax vs two-pass softmax


@pytest.mark.parametrize("B,heads,T", [(2, 3, 256), (3, 4, 64), (1, 2, 2048), (2, 4, 100), (1, 4, 512), (2, 1, 129)])
def test_attention_tensor_core_bf16(L, B, heads, T):
    """tcgen05 attention on plane-major bf16 qkv vs torch fp32 on the same bf16-rounded inputs.  P is rounded to
    bf16 before P V and exp2 is the approximate one: |err| <= 2e-2 on O(1) outputs, mean <= 2e-3."""
    from ccdm_b200 import _lib
    from ccdm_b200.engine import from_pm, to_pm
    D = 32
    C = heads * D
    qkv = (_rand(B, 3 * C, T, seed=63) * 1.3).to(torch.bfloat16).float()
    q, k, v = qkv.reshape(B * heads, 3 * D, T).split(D, dim=1)
    s = 1 / math.sqrt(math.sqrt(D))
    wgt = torch.softmax(torch.einsum("bct,bcs->bts", q * s, k * s), dim=-1)
    ref = torch.einsum("bts,bcs->bct", wgt, v).reshape(B, C, T)
    src = to_pm(qkv.permute(0, 2, 1).reshape(B, 1, T, 3 * C).to(torch.bfloat16).cuda())  # [B, 3C/8, 1, T, 8]
    out = torch.full((B, C // 8, 1, T, 8), float("nan"), dtype=torch.bfloat16, device="cuda")
    op = _lib.Op(kind=_lib.OP_ATTENTION, dtype=_lib.DT_BF16, out_dtype=_lib.DT_BF16, B=B, Hin=1, Win=T, Hout=1, Wout=T,
                 C0=3 * C, Cout=C, heads=heads, head_dim=D, exact=0)
    op.src0, op.out = src.data_ptr(), out.data_ptr()
    _lib.check(L.ccdm_launch_op(ctypes.byref(op), _sp()))
    torch.cuda.synchronize()
    got = from_pm(out).reshape(B, T, C).float().cpu().permute(0, 2, 1)
    err = (got - ref).abs()
    assert float(err.max()) < 2e-2 and float(err.mean()) < 2e-3, (float(err.max()), float(err.mean()))
    # the exact fp32-maths kernel on the same plane-major bf16 tensors (exact=1) agrees as well
    out2 = torch.full_like(out, float("nan"))
    op.exact, op.out = 1, out2.data_ptr()
    _lib.check(L.ccdm_launch_op(ctypes.byref(op), _sp()))
    torch.cuda.synchronize()
    err2 = (from_pm(out2).reshape(B, T, C).float().cpu().permute(0, 2, 1) - ref).abs()
    assert float(err2.max()) < 1.2e-2, float(err2.max())


# ---------------------------------------------------------------------------------------
# timestep-embedding table vs torch
# ---------------------------------------------------------------------------------------
def test_time_table(L):
    from ccdm_b200 import _lib
    from oracle.unet_ref import _timestep_embedding
    mc, cols = 32, 2016
    ed = 4 * mc
    w0, b0 = _rand(ed, mc, seed=71) / math.sqrt(mc), _rand(ed, seed=72) * 0.1
    w2, b2 = _rand(ed, ed, seed=73) / math.sqrt(ed), _rand(ed, seed=74) * 0.1
    wa, ba = _rand(cols, ed, seed=75) / math.sqrt(ed), _rand(cols, seed=76) * 0.1
    t = torch.tensor([1.0, 2.0, 37.0, 250.0, 999.0, 1000.0, 0.5])
    emb = F.linear(F.silu(F.linear(_timestep_embedding(t, mc), w0, b0)), w2, b2)
    ref = F.linear(F.silu(emb), wa, ba)
    dev = [x.cuda().contiguous() for x in (t, w0, b0, w2, b2, wa, ba)]
    out = torch.full((len(t), cols), float("nan"), device="cuda")
    _lib.check(L.ccdm_time_table(dev[0].data_ptr(), len(t), mc, *[d.data_ptr() for d in dev[1:]], cols, out.data_ptr(), _sp()))
    torch.cuda.synchronize()
    # cos/sin of arguments up to 1000 rad in fp32: 1 ulp of the argument is 6e-5
    assert float((out.cpu() - ref).abs().max()) < 3e-4


def test_nchw_to_nhwc_stats(L):
    from ccdm_b200 import _lib
    B, C, H, W = 2, 384, 8, 16
    x = _rand(B, C, H, W, seed=81).cuda()
    dst = torch.zeros((B, H, W, C), device="cuda")
    stat = torch.zeros((B, C, 2), dtype=torch.float64, device="cuda")
    _lib.check(L.ccdm_nchw_to_nhwc_stats(x.data_ptr(), B, C, H, W, _lib.DT_F32, dst.data_ptr(), stat.data_ptr(), 1, _sp()))
    torch.cuda.synchronize()
    assert torch.equal(dst, x.permute(0, 2, 3, 1).contiguous())
    xd = x.double()
    np.testing.assert_allclose(stat[..., 0].cpu().numpy(), xd.sum(dim=(2, 3)).cpu().numpy(), rtol=1e-12, atol=1e-9)
    np.testing.assert_allclose(stat[..., 1].cpu().numpy(), (xd * xd).sum(dim=(2, 3)).cpu().numpy(), rtol=1e-12, atol=1e-9)


def test_errors_are_loud(L):
    from ccdm_b200 import _lib
    op = _lib.Op(kind=_lib.OP_CONV, B=1, Hin=8, Win=8, Hout=8, Wout=8, C0=30, Cout=32, ksize=3, stride=1)
    assert L.ccdm_launch_op(ctypes.byref(op), _sp()) != 0
    assert b"multiples of 8" in L.ccdm_last_error()
    op = _lib.Op(kind=99)
    assert L.ccdm_launch_op(ctypes.byref(op), _sp()) != 0
