"""GPU parity of the DINO ViT condition encoder on the B200, through the C ABI.

Kernel level: every kernel of csrc/vit.cu against the same op in float64 / torch on the CPU.
  vit_linear (tcgen05, split fp16 operands): error <= LIN_REL = 5e-6 of the output's scale -- the tolerance the sampler's
  fp16x2 convs are held to (tests/test_gpu_exact.py); layernorm, patch embedding, position-embedding resize, descriptor map:
  <= 2e-6 of the scale plus the 2^-22 of the storage format.
Encoder level: ccdm_b200.models.condition_encoder.DinoViT against the fixtures the reference's own ViTExtractor produced
(tests/golden/make_golden_dino.py) and against the oracle: max |difference| <= DESC_TOL = 2e-4 on descriptors of unit
scale (|x| up to ~6) after up to 11 transformer blocks; measured values are printed.
"""
import math

import numpy as np
import pytest
import torch

from conftest import golden
from dino_cases import CASES, image

pytestmark = pytest.mark.gpu

LIN_REL = 5e-6
DESC_TOL = 2e-4


@pytest.fixture(scope="module")
def L(cuda_device):
    from ccdm_b200 import _lib
    _lib.require_device()
    return _lib.lib()


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


def _sp():
    from ccdm_b200 import _lib
    return _lib.stream_ptr(torch.cuda.current_stream())


def _shift(w):
    return int(max(0, min(13, math.floor(math.log2(32768.0 / float(w.abs().max()))))))


@pytest.mark.parametrize("B,T,Cin,Cout,gelu,res", [(2, 300, 384, 384, 0, 0), (1, 2049, 384, 1152, 0, 0), (2, 129, 384, 1536, 1, 0),
                                                     (1, 517, 1536, 384, 0, 1), (3, 1, 384, 384, 0, 1), (1, 128, 768, 768, 1, 1),
                                                     (2, 65, 32, 128, 0, 0)])
def test_vit_linear(L, B, T, Cin, Cout, gelu, res):
    from ccdm_b200 import _lib
    from ccdm_b200.engine import pack_conv_weight_x3
    from ccdm_b200.vit_engine import tokens_from_float, tokens_to_float
    x = _rand(B, T, Cin, seed=1) * 1.3 + 0.1
    w = _rand(Cout, Cin, seed=2) / math.sqrt(Cin)
    b = _rand(Cout, seed=3) * 0.2
    r = _rand(B, T, Cout, seed=4) * 2.0 if res else None
    shift = _shift(w)
    xd = tokens_from_float(x).cuda()
    x_st = tokens_to_float(xd, B, Cin, T).cpu().double()  # what the kernel actually reads
    wp = pack_conv_weight_x3(w.reshape(Cout, Cin, 1, 1), shift, int(L.ccdm_vit_linear_nt())).cuda().contiguous()
    bd = b.cuda()
    rd = tokens_from_float(r).cuda() if res else None
    out = torch.empty(B * Cout * T * 2, dtype=torch.float16, device="cuda")
    _lib.check(L.ccdm_vit_linear(xd.data_ptr(), wp.data_ptr(), bd.data_ptr(), rd.data_ptr() if res else None, B, T, Cin, Cout, gelu,
                                 shift + _lib.F16X2_SCALE_LOG2, out.data_ptr(), _sp()), "vit_linear")
    torch.cuda.synchronize()
    got = tokens_to_float(out, B, Cout, T).cpu().double()
    ref = x_st @ w.double().t() + b.double()
    if gelu:
        ref = torch.nn.functional.gelu(ref)
    if res:
        ref = ref + tokens_to_float(rd, B, Cout, T).cpu().double()
    scale = float(ref.abs().max())
    err = float((got - ref).abs().max()) / scale
    print(f"vit_linear B{B} T{T} {Cin}->{Cout} gelu{gelu} res{res}: rel err {err:.2e}")
    assert err <= LIN_REL


def test_vit_linear_rejects_bad_shapes(L):
    x = torch.zeros(64, dtype=torch.float16, device="cuda")
    f = torch.zeros(64, dtype=torch.float32, device="cuda")
    assert L.ccdm_vit_linear(x.data_ptr(), x.data_ptr(), f.data_ptr(), None, 1, 1, 24, 128, 0, 10, x.data_ptr() + 16, _sp()) < 0   # Cin % 32
    assert L.ccdm_vit_linear(x.data_ptr(), x.data_ptr(), f.data_ptr(), None, 1, 1, 32, 96, 0, 10, x.data_ptr() + 16, _sp()) < 0    # Cout % 128
    assert L.ccdm_vit_linear(x.data_ptr(), x.data_ptr(), f.data_ptr(), None, 1, 1, 32, 128, 0, 10, x.data_ptr(), _sp()) < 0        # aliasing
    assert b"alias" in L.ccdm_last_error()


@pytest.mark.parametrize("B,T,C", [(2, 2049, 384), (1, 33, 768), (3, 1, 384), (1, 100, 1024)])
def test_vit_layernorm(L, B, T, C):
    from ccdm_b200 import _lib
    from ccdm_b200.vit_engine import tokens_from_float, tokens_to_float
    x = _rand(B, T, C, seed=5) * 2.0 + 0.7
    g = 1.0 + 0.1 * _rand(C, seed=6)
    be = 0.1 * _rand(C, seed=7)
    xd = tokens_from_float(x).cuda()
    x_st = tokens_to_float(xd, B, C, T).cpu().double()
    out = torch.empty_like(xd)
    gd, bd = g.cuda(), be.cuda()
    _lib.check(L.ccdm_vit_layernorm(xd.data_ptr(), gd.data_ptr(), bd.data_ptr(), B, T, C, 1e-6, out.data_ptr(), _sp()), "ln")
    torch.cuda.synchronize()
    got = tokens_to_float(out, B, C, T).cpu().double()
    ref = torch.nn.functional.layer_norm(x_st, (C,), g.double(), be.double(), 1e-6)
    err = float((got - ref).abs().max()) / float(ref.abs().max())
    print(f"vit_layernorm B{B} T{T} C{C}: rel err {err:.2e}")
    assert err <= 2e-6


@pytest.mark.parametrize("B,H,W,p,s,D", [(2, 64, 128, 8, 8, 384), (1, 72, 104, 8, 4, 384), (1, 48, 80, 16, 16, 768), (1, 40, 56, 8, 8, 768)])
def test_vit_patch_embed_and_pos_embed(L, B, H, W, p, s, D):
    from ccdm_b200 import _lib
    from ccdm_b200.vit_engine import tokens_to_float
    from oracle import dino_ref
    img = _rand(B, 3, H, W, seed=8)
    w = _rand(D, 3, p, p, seed=9) / math.sqrt(3 * p * p)
    b = _rand(D, seed=10) * 0.1
    cls = _rand(D, seed=11) * 0.2
    n = 224 // p
    pos = _rand(1 + n * n, D, seed=12) * 0.2
    hp, wp = 1 + (H - p) // s, 1 + (W - p) // s
    T = 1 + hp * wp
    # position embedding: F.interpolate(bicubic) through the oracle's restatement of dino.py:86-117
    holder = type("M", (), {})()
    holder.pos_embed = pos[None]
    want_pos = dino_ref.fix_pos_enc(holder, T, H, W, p, (s, s))[0]
    pos_d = torch.empty(T, D, device="cuda")
    pos_src, img_d, b_d, cls_d = pos.cuda(), img.cuda(), b.cuda(), cls.cuda()
    _lib.check(L.ccdm_vit_pos_embed(pos_src.data_ptr(), n, D, hp, wp, (hp + 0.1) / n, (wp + 0.1) / n, pos_d.data_ptr(), _sp()), "pos")
    torch.cuda.synchronize()
    e_pos = float((pos_d.cpu() - want_pos).abs().max())
    print(f"vit_pos_embed {n}x{n} -> {hp}x{wp}: max err {e_pos:.2e}")
    assert e_pos <= 2e-6
    out = torch.empty(B * D * T * 2, dtype=torch.float16, device="cuda")
    wt = w.reshape(D, -1).t().contiguous().cuda()
    _lib.check(L.ccdm_vit_patch_embed(img_d.data_ptr(), wt.data_ptr(), b_d.data_ptr(), cls_d.data_ptr(), pos_d.data_ptr(),
                                      B, H, W, p, s, D, out.data_ptr(), _sp()), "patch_embed")
    torch.cuda.synchronize()
    got = tokens_to_float(out, B, D, T).cpu().double()
    tok = torch.nn.functional.conv2d(img.double(), w.double(), b.double(), stride=s).flatten(2).transpose(1, 2)
    ref = torch.cat([cls.double().expand(B, 1, D), tok], 1) + pos_d.cpu().double()[None]
    err = float((got - ref).abs().max()) / float(ref.abs().max())
    print(f"vit_patch_embed {H}x{W} p{p} s{s} D{D}: rel err {err:.2e}")
    assert err <= 2e-6


@pytest.mark.parametrize("B,hp,wp,Ho,Wo,heads", [(2, 8, 16, 8, 16, 6), (1, 17, 25, 18, 26, 6), (1, 12, 8, 20, 12, 12), (1, 5, 7, 3, 4, 6)])
def test_vit_descriptor(L, B, hp, wp, Ho, Wo, heads):
    from ccdm_b200 import _lib
    from ccdm_b200.vit_engine import tokens_from_float, tokens_to_float
    hd = 64
    C, T = heads * hd, 1 + hp * wp
    k = _rand(B, T, C, seed=13)
    kd = tokens_from_float(k).cuda()
    k_st = tokens_to_float(kd, B, C, T).cpu()
    out = torch.empty(B, C, Ho, Wo, device="cuda")
    _lib.check(L.ccdm_vit_descriptor(kd.data_ptr(), B, T, heads, hd, hp, wp, Ho, Wo, out.data_ptr(), _sp()), "descriptor")
    torch.cuda.synchronize()
    x = k_st.reshape(B, T, heads, hd).permute(0, 2, 1, 3)[:, :, 1:, :]            # [B, heads, t, d] without cls (dino.py:293-297)
    x = x.permute(0, 2, 3, 1).flatten(start_dim=-2, end_dim=-1).view(B, hp, wp, -1).permute(0, 3, 1, 2)
    ref = torch.nn.functional.interpolate(x, (Ho, Wo), mode="bilinear")
    err = float((out.cpu() - ref).abs().max())
    print(f"vit_descriptor {hp}x{wp} -> {Ho}x{Wo}: max err {err:.2e}")
    assert err <= (0.0 if (Ho, Wo) == (hp, wp) else 2e-6)


def _encoder(mt, stride, seed=0):
    from ccdm_b200.models.condition_encoder import DinoViT
    from ccdm_b200.synthetic import fill_synthetic_
    enc = DinoViT(mt, False, "concat_pixels_concat_features", stride=stride)
    fill_synthetic_(enc.extractor.model, seed)
    return enc.to("cuda").eval()


def _golden_err(out, g):
    assert tuple(out.shape) == tuple(g["shape"])
    if "desc" in g:
        return float(np.abs(out - g["desc"]).max())
    h0, w0 = (int(v) for v in g["win0"])
    return float(max(np.abs(out[:, :, h0:h0 + 8, w0:w0 + 8] - g["window"]).max(),
                     np.abs(out.astype(np.float64).mean(axis=(2, 3)) - g["chan_mean"]).max(),
                     np.abs(out[:, ::7, ::3, ::5] - g["strided"]).max()))


@pytest.mark.parametrize("tag", sorted(CASES))
def test_encoder_reproduces_reference_fixture(cuda_device, tag):
    """DinoViT.forward == the reference's ViTExtractor.extract_descriptors on the same weights and images."""
    mt, stride, B, H, W, layers, rs = CASES[tag]
    enc = _encoder(mt, stride)
    enc.layers, enc.resize_shape = layers, rs
    x = image(B, H, W, 77).cuda()
    out = enc(x)
    assert out.dtype == torch.float32 and out.is_cuda
    err = _golden_err(out.cpu().numpy(), golden(tag + ".npz"))
    print(f"encoder {tag}: max |diff| vs the reference fixture {err:.2e}")
    assert err <= DESC_TOL


def test_encoder_blockwise_against_oracle(cuda_device):
    """Token stream after the embedding and after every block, against the oracle run on the same image: the error stays
    at fp32 level through the depth (no drift), and a multi-layer call returns each layer's descriptors."""
    from ccdm_b200.vit_engine import tokens_to_float
    from ccdm_b200.synthetic import fill_synthetic_
    from oracle import dino_ref
    B, H, W = 1, 64, 96
    enc = _encoder("dino_vits8", 8, seed=2)
    x = image(B, H, W, 31)
    toks = {}
    outs = enc.extractor.engine().key_descriptors(x.cuda(), [3, 11], [None, None], tokens_out=toks)
    torch.cuda.synchronize()
    vit = fill_synthetic_(dino_ref.build("dino_vits8"), 2).eval()
    T = 1 + (H // 8) * (W // 8)
    with torch.no_grad():
        t = vit.prepare_tokens(x)
        e0 = float((tokens_to_float(toks["tokens"], B, 384, T).cpu() - t).abs().max())
        worst = e0
        for i in range(11):
            t = vit.blocks[i](t)
            e = float((tokens_to_float(toks["block%d" % i], B, 384, T).cpu() - t).abs().max()) / float(t.abs().max())
            worst = max(worst, e)
        want = dino_ref.extract_descriptors(vit, x, [3, 11], 8, None)
    print(f"encoder blockwise: embedding err {e0:.2e}, worst relative token err over 11 blocks {worst:.2e}")
    assert worst <= 2e-5
    for a, b in zip(outs, want):
        assert float((a.cpu() - b).abs().max()) <= DESC_TOL


def test_encoder_drop_in_with_the_sampler(cuda_device):
    """Evaluator.predict_feature_condition + predict (eval_cdm.py:154-174): the encoder's output is the feature_condition of a
    DINO-conditioned DenoisingModel; a checkpoint entry of the encoder loads strictly and takes effect (weight cache)."""
    from conftest import DINO, build_ours
    from ccdm_b200.synthetic import fill_synthetic_, synthetic_inputs
    from oracle import dino_ref
    B, H, W, K = 2, 64, 128, 20
    enc = _encoder("dino_vits8", 8)
    model = build_ours(250, 3, H, W, K, "majority", DINO, (1, 1, 2, 2, 4, 4)).cuda().eval()
    img, _, labels = synthetic_inputs(B, 3, H, W, K)
    with torch.no_grad():
        enc.eval()
        feat = enc(img.cuda())                                   # predict_feature_condition
        assert tuple(feat.shape) == (B, 384, H // 8, W // 8)
        xt = torch.nn.functional.one_hot(labels.long(), K).permute(0, 3, 1, 2).float().cuda()
        model.noise, model.seed = "philox", 1
        ret = model(x=xt, condition=img.cuda(), feature_condition=feat, t=torch.as_tensor(10003))
        assert ret["diffusion_out"].shape == (B, K, H, W)
    # a "feature_cond_encoder" checkpoint entry (eval_cdm.py:136-142): strict load, then the packed weights follow
    other = fill_synthetic_(dino_ref.build("dino_vits8"), 9)
    ckpt = {"extractor.model." + k: v for k, v in other.state_dict().items()}
    enc.load_state_dict(ckpt, strict=True)
    with torch.no_grad():
        feat2 = enc(img.cuda())
        want = dino_ref.extract_descriptors(other.eval(), img, 11, 8, None)
    assert float((feat2.cpu() - want).abs().max()) <= DESC_TOL
    assert float((feat2 - feat).abs().max()) > 0.1
