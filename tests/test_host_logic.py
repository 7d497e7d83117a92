"""Host-side logic, CPU only: the reference-shaped surface, program planning, the C ABI."""
import ctypes
import json
import os
import re

import numpy as np
import pytest
import torch

from conftest import DINO, GOLDEN, REFERENCE, ROOT, UNET_PARAMS, build_ours, golden


def _build(C, H, W, K, fce, mult, base=32, **kw):
    from ccdm_b200 import models
    p = dict(UNET_PARAMS, channel_mult=mult, base_channels=base)
    p.update(kw)
    return models.build_model(250, "cosine", {"s": 0.008}, [(C, H, W), (K, H, W)], (C, H, W), "unet_openai", p, "d",
                              "majority", DINO if fce else None)


def test_state_dict_layout_matches_reference_manifest():
    """Key names, order and shapes of state_dict() for five configurations, recorded from the reference."""
    man = json.load(open(os.path.join(GOLDEN, "state_dict_manifest.json")))
    for tag, rec in man.items():
        C, H, W, K, fce, mult, base = rec["args"]
        m = _build(C, H, W, K, fce, tuple(mult) if mult else None, base)
        got = [[k, list(v.shape)] for k, v in m.state_dict().items()]
        assert got == rec["keys"], tag


def test_schedule_buffers_bit_identical_to_reference():
    from ccdm_b200.models import DiffusionModel
    g = golden("schedules.npz")
    for T in (100, 250, 1000):
        d = DiffusionModel("cosine", T, 2, {"s": 0.008})
        for n in ("betas", "alphas", "cumalphas"):
            np.testing.assert_array_equal(getattr(d, n).numpy(), g[f"cosine{T}_{n}"])
    d = DiffusionModel("linear", 250, 2, None)
    for n in ("betas", "alphas", "cumalphas"):
        np.testing.assert_array_equal(getattr(d, n).numpy(), g[f"linear250_{n}"])
    assert d.time_steps == 250
    assert d.step_scalars(1) == (0.0, 1.0)
    assert d.step_scalars(2) == (float(d.alphas[1]), float(d.cumalphas[0]))


def test_reverse_t_values_match_reference_and_oracle():
    from ccdm_b200.models.diffusion_denoising import reverse_t_values
    from oracle import cdm
    g = golden("t_values.npz")
    for key in g.files:
        T, req = key[1:].split("_req")
        req = None if req == "None" else int(req)
        assert reverse_t_values(int(T), req) == g[key].tolist() == cdm.t_values(int(T), req)
    with pytest.raises(AssertionError):
        reverse_t_values(250, 10000 + 300)


def test_constructor_surface_and_errors():
    from ccdm_b200 import models
    from ccdm_b200.models.unet_openai import create_unet_openai
    with pytest.raises(NotImplementedError, match="backbone resnet50"):
        models.build_model(250, "cosine", None, [(1, 64, 64), (2, 64, 64)], None, "resnet50", {}, "d")
    with pytest.raises(ValueError, match="unsupported image size"):
        create_unet_openai(100, 32, 3, 2, 2, None)
    for flag in ("use_fp16", "use_scale_shift_norm", "resblock_updown", "use_new_attention_order", "ce_head"):
        with pytest.raises(NotImplementedError):
            create_unet_openai(64, 32, 3, 2, 2, None, **{flag: True})
    with pytest.raises(KeyError):
        models.DiffusionModel("quadratic", 10, 2)
    m = _build(1, 64, 64, 2, False, None)
    assert m.unet.in_channels == 3 and m.unet.out_channels == 2 and m.unet.model_channels == 32
    assert m.unet.channel_mult == (1, 2, 3, 4) and m.time_steps == 250 and m.diffusion.num_classes == 2
    assert m.step_T_sample == "majority" and m.dataset_file == "d" and m.unet.feature_condition_idx == []
    assert _build(3, 256, 512, 20, True, None).unet.feature_condition_idx == [10]
    # fresh constructor = the reference's degenerate zero-initialised heads (unet.py:216-218,300,705)
    sd = m.unet.state_dict()
    assert float(sd["out.2.weight"].abs().sum()) == 0 and float(sd["input_blocks.1.0.out_layers.3.weight"].abs().sum()) == 0
    assert float(sd["middle_block.1.proj_out.weight"].abs().sum()) == 0
    assert float(sd["out.0.weight"].min()) == 1.0


def test_forward_dispatch_errors_without_gpu():
    m = build_ours(250, 1, 64, 64, 2)
    x = torch.zeros(1, 2, 64, 64)
    m.train()
    with pytest.raises(ValueError, match="'t' needs to be a Tensor"):
        m(x, torch.zeros(1, 1, 64, 64), None, None)
    m.eval()
    with pytest.raises(NotImplementedError):
        m(x, torch.zeros(1, 1, 64, 64), None, label_ref_logits=torch.zeros(1))
    if not torch.cuda.is_available():
        from ccdm_b200 import _lib
        with pytest.raises(_lib.CcdmError):  # no CPU fallback
            m(x, torch.zeros(1, 1, 64, 64), None)
    from ccdm_b200 import _lib
    with pytest.raises(_lib.CcdmError):  # the device-side x_T draw needs the model on a GPU, too
        m.draw_x_T(1, 64, 64)


def test_onehot_categorical_matches_torch_distributions():
    from ccdm_b200.models import OneHotCategoricalBCHW
    g = golden("draw.npz")
    for K in (2, 20):
        p = torch.as_tensor(g[f"K{K}_probs_in"]).permute(0, 3, 1, 2).contiguous()  # NCHW memory, as in the fixture
        torch.manual_seed(1000 + K)
        s = OneHotCategoricalBCHW(probs=p).sample()
        assert s.dtype == torch.float32 and not s.is_contiguous()
        np.testing.assert_array_equal(s.argmax(1).numpy(), g[f"K{K}_sample_labels"])
        d = OneHotCategoricalBCHW(probs=p)
        assert d.max_prob_sample().dtype == torch.int64
        np.testing.assert_array_equal(d.max_prob_sample().argmax(1).numpy(), g[f"K{K}_majority_labels"])
        np.testing.assert_array_equal(d.prob_sample().permute(0, 2, 3, 1).numpy(), g[f"K{K}_confidence"])
        torch.manual_seed(2000 + K)
        xT = OneHotCategoricalBCHW(logits=torch.zeros(p.shape)).sample()
        np.testing.assert_array_equal(xT.argmax(1).numpy(), g[f"K{K}_xT_labels"])
    with pytest.raises(ValueError):
        OneHotCategoricalBCHW(probs=torch.ones(3))


def test_c_abi_exports_every_declared_symbol():
    from ccdm_b200 import _lib
    header = open(os.path.join(ROOT, "include", "ccdm_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    names = set(re.findall(r"\b(ccdm_[a-z0-9_]+)\s*\(", header))
    assert len(names) >= 18
    L = _lib.lib()  # also checks ABI version and struct sizes against the ctypes mirrors
    for n in sorted(names):
        assert hasattr(L, n), f"{n} declared in include/ccdm_b200.h but not exported"
    assert L.ccdm_abi_version() == _lib.ABI_VERSION
    assert ctypes.sizeof(_lib.StepEntry) == 32


@pytest.mark.parametrize("cfg", [(1, 128, 128, 2, False, 3), (3, 256, 512, 20, True, 1), (1, 64, 64, 2, False, 2)])
def test_program_planning_dry_run(cfg):
    """Op list, liveness-planned workspace and struct marshalling, without a GPU."""
    from ccdm_b200 import _lib
    C, H, W, K, fce, B = cfg
    m = _build(C, H, W, K, fce, None)
    eng = m.unet.engine("fp32", dry_run=True)
    eng.weights.refresh()
    prog = eng.program(B, H, W)
    prog.bind(250)
    ops = prog._op_array
    kinds = [o.kind for o in ops]
    assert kinds[0] == _lib.OP_INPUT_CONV and kinds[-1] == _lib.OP_HEAD and kinds[-2] == _lib.OP_CONV
    n_res = sum(1 for b in m.unet.arch.blocks for l in b.layers if l.kind == "res")
    n_att = sum(1 for b in m.unet.arch.blocks for l in b.layers if l.kind == "attn")
    n_updown = sum(1 for b in m.unet.arch.blocks for l in b.layers if l.kind in ("up", "down"))
    assert len(ops) == 1 + 2 * n_res + 3 * n_att + n_updown + 2
    assert kinds.count(_lib.OP_ATTENTION) == n_att
    # no live tensor overlaps another one it coexists with
    tens = sorted(prog.tens.values(), key=lambda t: t.off)
    for i, a in enumerate(tens):
        for b in tens[i + 1:]:
            if b.off >= a.off + a.nbytes:
                break
            assert a.last < b.born or b.last < a.born, (a.name, b.name)
    # every op reads tensors that are alive and writes inside the arena
    lo, hi = prog.addr["arena"], prog.addr["arena"] + prog.arena_bytes
    for o in ops:
        if o.kind in (_lib.OP_CONV, _lib.OP_INPUT_CONV, _lib.OP_ATTENTION):
            assert lo <= o.out < hi
            assert o.B == B and o.Cout > 0
        if o.kind == _lib.OP_CONV and o.gn:
            assert o.stat0 and o.gamma and o.beta and ((o.C0 + o.C1) % 32 == 0 or o.gn_cpg > 0)
    # concat never materialised: output blocks read two sources
    assert sum(1 for o in ops if o.kind == _lib.OP_CONV and o.C1 > 0) == sum(1 for b in m.unet.arch.blocks if b.stage == "out" or b.feat_concat)
    assert sum(1 for o in ops if o.S0 > 0) == sum(1 for b in m.unet.arch.blocks for l in b.layers if l.kind == "res" and l.skip_conv)
    # reuse actually happens
    assert prog.arena_bytes < 0.6 * sum(t.nbytes for t in prog.tens.values())
    with pytest.raises(_lib.CcdmError):
        eng.run_chain(torch.zeros(B, K, H, W), torch.zeros(B, C, H, W), None, [1], [0.0], [1.0], 1)


def test_weight_packing_layout():
    from ccdm_b200.engine import pack_bias, pack_conv_weight
    w = torch.arange(2 * 3 * 3 * 3, dtype=torch.float32).reshape(2, 3, 3, 3)
    p = pack_conv_weight(w, 8)
    assert p.shape == (9, 8, 32)
    assert p[4, 1, 1] == w[1, 1, 1, 1] and p[2, 2, 0] == w[0, 2, 0, 2] and float(p[:, 3:, :].abs().sum()) == 0
    assert float(p[:, :, 2:].abs().sum()) == 0
    q = pack_conv_weight(torch.ones(96, 32, 1))
    assert q.shape == (1, 32, 96)
    assert pack_bias(torch.ones(20)).shape == (32,)
    m = build_ours(250, 1, 64, 64, 2)
    eng = m.unet.engine("fp32", dry_run=True)
    assert eng.weights.refresh() is True and eng.weights.refresh() is False
    with torch.no_grad():
        m.unet.time_embed._modules["0"].bias.add_(1.0)
    assert eng.weights.refresh() is True  # in-place writes are seen (polyak.py:18-26 hazard)
    sd = m.unet.state_dict()
    np.testing.assert_array_equal(eng.weights.view("te0_b").numpy(), sd["time_embed.0.bias"].numpy())
    cols = m.unet.arch.emb_cols
    first = next(l for b in m.unet.arch.blocks for l in b.layers if l.kind == "res")
    np.testing.assert_array_equal(eng.weights.view("emb_w")[:first.cout * 128].reshape(first.cout, 128).numpy(),
                                  sd[first.path + ".emb_layers.1.weight"].numpy())
    assert cols == sum(l.cout for b in m.unet.arch.blocks for l in b.layers if l.kind == "res")


def test_tensor_core_weight_packing_and_n_tile_rule():
    """[cc][Cin/8][tap][NT][8] packing of the tcgen05 kernels: NT is a function of Cout only (one packed tensor serves
    every shape), <= 64 up to 128 output channels, <= 192 for the wider q/k/v layers; every weight lands where the
    kernel's B-operand descriptor reads it."""
    from ccdm_b200 import _lib
    from ccdm_b200.engine import pack_conv_weight_tc
    L = _lib.lib()
    want = {2: 16, 20: 32, 32: 32, 64: 64, 96: 48, 128: 64, 192: 192, 288: 144, 384: 192, 160: 160, 224: 112}
    for cout, nt in want.items():
        got = int(L.ccdm_conv_tc_nt(cout, 1, 0))
        cop = (cout + 15) // 16 * 16
        assert got == nt and cop % got == 0 and got % 16 == 0 and got <= 192, (cout, got)
        got9 = int(L.ccdm_conv_tc_nt(cout, 9, 0))  # 3x3 (and sub-pixel) convs: never wider than 64, so a weight stage stays small
        assert got9 == (nt if cout <= 128 else max(d for d in range(16, 65, 16) if cop % d == 0)) and cop % got9 == 0
    g = torch.Generator().manual_seed(5)
    for (co, ci, k) in [(384, 128, 1), (96, 32, 3), (20, 32, 3)]:
        w = torch.randn(co, ci, k, k, generator=g)
        pk = pack_conv_weight_tc(w)
        nt = int(L.ccdm_conv_tc_nt(co, k * k, 0))
        cop = (co + 15) // 16 * 16
        assert pk.dtype == torch.bfloat16 and pk.shape == (cop // nt, ci // 8, k * k, nt, 8)
        wb = w.to(torch.bfloat16)
        for (o, c, t) in [(0, 0, 0), (co - 1, ci - 1, k * k - 1), (co // 2, 9 % ci, (k * k) // 2)]:
            assert pk[o // nt, c // 8, t, o % nt, c % 8] == wb[o, c, t // k, t % k]
        if cop > co:  # padded output channels are zero rows of the B operand
            assert float(pk.float().reshape(cop // nt, ci // 8, k * k, nt, 8)[-1, :, :, co % nt:, :].abs().sum()) == 0.0


def test_tensor_core_conv_configuration_rules():
    """Host-side tile / pipeline choices of the TMA-fed tcgen05 conv kernel (no GPU needed): K chunks of 16 channels on
    32-channel inputs and on the mid-size levels, 32 elsewhere; persistent grid of at most one CTA per SM; the whole
    configuration fits the 227 KB of shared memory and the 512 TMEM columns."""
    import ctypes
    from ccdm_b200 import _lib
    L = _lib.lib()

    def cfg(B, H, W, C0, C1, Cout, k=3, S0=0, S1=0):
        op = _lib.Op(kind=_lib.OP_CONV, dtype=_lib.DT_BF16, out_dtype=_lib.DT_BF16, B=B, Hin=H, Win=W, Hout=H, Wout=W, C0=C0, C1=C1,
                     Cout=Cout, ksize=k, stride=1, S0=S0, S1=S1, gn=1, silu=1)
        out = (ctypes.c_int32 * 16)()
        assert L.ccdm_conv_tc_config(ctypes.byref(op), out) == 0 and L.ccdm_conv_uses_tma(ctypes.byref(op)) == 1
        names = ["PL", "R", "Wt", "MB", "NQ", "NT", "n_cc", "NS", "resident", "acc2", "tmem", "tiles", "items", "grid", "smem", "chunks"]
        return dict(zip(names, out))

    full = cfg(64, 128, 128, 32, 32, 32)          # LIDC out12-14: concat 64 -> 32 at full resolution
    assert full["PL"] == 4 and full["chunks"] == 2
    assert cfg(64, 128, 128, 32, 0, 32)["PL"] == 2  # 32-channel input: two chunks of 16 so the roles of one item overlap
    mid = cfg(8, 128, 256, 32, 32, 32)              # Cityscapes 128x256 at B = 8: mid-size level -> chunks of 16
    assert mid["PL"] == 2 and mid["chunks"] == 4 and mid["R"] > cfg(8, 256, 512, 32, 32, 32)["R"]
    for c in (full, mid, cfg(64, 8, 8, 128, 0, 128, S0=128), cfg(64, 8, 8, 128, 0, 384, k=1), cfg(8, 32, 64, 64, 384, 64)):
        assert 1 <= c["grid"] <= 148 and c["items"] >= c["grid"]
        assert c["smem"] <= 227 * 1024 and c["tmem"] in (32, 64, 128, 256, 512)
        assert c["MB"] * c["NT"] * (2 if c["acc2"] else 1) <= 512 and c["NS"] >= 2


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference checkout not present (GPU box)")
def test_live_reference_agrees_with_oracle_on_fresh_seed():
    """Beyond the committed fixtures: a different seed / shape, reference imported live."""
    import sys
    sys.path.insert(0, REFERENCE)
    import models as ref_models
    from ccdm_b200.synthetic import fill_synthetic_, synthetic_inputs
    from oracle import chain_ref, unet_ref
    sys.path.remove(REFERENCE)
    p = dict(UNET_PARAMS)
    torch.manual_seed(3)
    ref = ref_models.build_model(100, "cosine", {"s": 0.008}, [(1, 64, 64), (2, 64, 64)], (1, 64, 64), "unet_openai", p,
                                 "datasets.lidc", "confidence", None).eval()
    fill_synthetic_(ref.unet, 5)
    ours = build_ours(100, 1, 64, 64, 2, "confidence", seed=5)
    for (k, a), (k2, b) in zip(ref.state_dict().items(), ours.state_dict().items()):
        assert k == k2 and torch.equal(a, b)
    image, _, labels = synthetic_inputs(1, 1, 64, 64, 2, seed=99)
    x = chain_ref.labels_to_onehot(labels.numpy(), 2)
    torch.manual_seed(8)
    with torch.no_grad():
        out = ref(x, image, None, t=torch.as_tensor(10005))["diffusion_out"]
    torch.manual_seed(8)
    _, probs = chain_ref.reverse_chain(ours.unet.state_dict(), labels.numpy(), image, None, ours.diffusion.alphas.numpy(),
                                       ours.diffusion.cumalphas.numpy(), 100, 10005, "confidence", K=2)
    np.testing.assert_allclose(probs, out.permute(0, 2, 3, 1).numpy(), rtol=0, atol=2e-6)


def test_subpixel_weights_reproduce_upsample_conv():
    """nearest x2 + conv3x3 == four 2x2 convs on the low-resolution input with pre-summed taps (Upsample, unet.py:106-116)."""
    import torch.nn.functional as F
    from ccdm_b200.engine import subpixel_weights
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 5, 7, 6, generator=g)
    w = torch.randn(4, 5, 3, 3, generator=g)
    ref = F.conv2d(F.interpolate(x, scale_factor=2, mode="nearest"), w, padding=1)
    wc = subpixel_weights(w)
    xp = F.pad(x, (1, 1, 1, 1))
    out = torch.zeros_like(ref)
    for py in (0, 1):
        for px in (0, 1):
            acc = 0
            for ry in (0, 1):
                for rx in (0, 1):
                    win = xp[:, :, py + ry:py + ry + 7, px + rx:px + rx + 6]
                    acc = acc + torch.einsum("bchw,oc->bohw", win, wc[:, :, 2 * py + px, 2 * ry + rx])
            out[:, :, py::2, px::2] = acc
    assert float((out - ref).abs().max()) < 1e-4


def test_every_conv_lands_on_the_tensor_core_kernel():
    """Dry-run plans (no GPU) of the shipped architectures -- base_channels 32 AND 64, LIDC 128x128 and Cityscapes 256x512
    with the DINO concat -- in both tensor-core modes: every conv-shaped op is taken by conv_tma (so its packed weights are
    the layout the kernel reads), no FFMA op is handed deferred-fold statistics rows, and the statistics layout a GroupNorm
    consumer is told matches what its producer writes."""
    import ctypes
    from ccdm_b200 import _lib, models
    from ccdm_b200.synthetic import fill_synthetic_
    L = _lib.lib()
    for prec in ("bf16", "exact"):
        for base in (32, 64):
            for (C, H, W, K, fce, B) in ((1, 128, 128, 2, False, 64), (3, 256, 512, 20, True, 8), (1, 64, 64, 2, False, 3)):
                m = _build(C, H, W, K, fce, None, base).eval()
                fill_synthetic_(m.unet, 0)
                eng = m.unet.engine(prec, dry_run=True)
                eng.weights.refresh()
                prog = eng.program(B, H, W)
                prog.bind(4)
                ops = prog._op_array
                n_conv = sum(1 for o in prog._op_dicts if o["kind"] == _lib.OP_CONV)
                assert prog.n_tc == n_conv and not prog.off_tc, (prec, base, H, W, prog.off_tc)
                producers = {}
                for i, o in enumerate(prog._op_dicts):
                    op = ops[i]
                    if op.kind == _lib.OP_CONV:
                        assert L.ccdm_conv_uses_tc(ctypes.byref(op)) == 1 and op.exact == 0
                        if o.get("_ws"):
                            # a fused skip chunk (1x1 skip conv / identity residual) is packed for the N tile of the conv it
                            # rides on, not for the tile a stand-alone 1x1 conv of its shape would get (they differ above 128
                            # output channels: found by tests/test_gpu_variants.py at base_channels = 64)
                            cfg = (ctypes.c_int32 * 16)()
                            assert L.ccdm_conv_tc_config(ctypes.byref(op), cfg) == 0
                            taps = 9 if op.ksize == 3 else 1
                            assert cfg[5] == L.ccdm_conv_tc_nt(op.Cout, taps, 1 if prec == "exact" else 0)
                            if o["_ws"].startswith("ident:"):
                                assert int(o["_ws"].split(":")[2]) == cfg[5], (o["_ws"], cfg[5])
                            else:
                                assert taps == 9  # ResBlock skip convs are always fused into the block's second 3x3 conv
                        for si, sten in enumerate(o.get("_src", [])[:2]):
                            slots = getattr(op, "st_slots%d" % si)
                            if o.get("gn") and sten.stat_layout is not None:
                                want = producers[id(sten)]
                                got = tuple(getattr(op, "%s%d" % (n, si)) for n in ("st_slots", "st_ips", "st_items", "st_grid", "st_rows"))
                                assert got == want and getattr(op, "stat%d" % si) == sten.part_addr
                            else:
                                assert slots == 0
                        if o["_out"] is not None and o["_out"].stat_layout is not None:
                            lay = (ctypes.c_int32 * 5)()
                            assert L.ccdm_conv_stat_layout(ctypes.byref(op), lay) == 0
                            producers[id(o["_out"])] = tuple(int(v) for v in lay)
                            assert op.ostat == 0 and op.part == o["_out"].part_addr
                if prec == "exact":
                    assert all(ops[i].acc_shift == eng.weights.shift + 4 for i in range(prog.n_ops) if ops[i].kind == _lib.OP_CONV)


def test_unsupported_head_dim_fails_at_plan_time():
    """An attention head size other than 32 or 64 channels is refused when the step program is built -- before any launch."""
    from ccdm_b200 import models
    p = dict(UNET_PARAMS, num_head_channels=-1, num_heads=1)   # one head of C = 96 / 128 channels
    m = models.build_model(50, "cosine", {"s": 0.008}, [(1, 64, 64), (2, 64, 64)], (1, 64, 64), "unet_openai", p, "datasets.lidc",
                           "majority", None).eval()
    for prec in ("exact", "bf16", "fp32"):
        eng = m.unet.engine(prec, dry_run=True)
        eng.weights.refresh()
        with pytest.raises(NotImplementedError, match="head_dim"):
            eng.program(2, 64, 64)
    p = dict(UNET_PARAMS, base_channels=64, num_head_channels=64)
    m = models.build_model(50, "cosine", {"s": 0.008}, [(1, 64, 64), (2, 64, 64)], (1, 64, 64), "unet_openai", p, "datasets.lidc", "majority", None).eval()
    for prec in ("exact", "bf16", "fp32"):
        eng = m.unet.engine(prec, dry_run=True)
        eng.weights.refresh()
        assert eng.program(2, 64, 64).n_ops > 0


@pytest.mark.parametrize("c_h,f", [(64, 384), (64, 768), (128, 384)])
def test_feature_fold_algebra(c_h, f):
    """conv3x3(SiLU(GN32(cat[h, f]))) == per-step conv over cat[h, f[:f16]] (groups of cpg channels, the feature
    channels past the shared group zero-weighted) + per-chain conv over f (same groups, counted from channel c_h; the shared
    group's feature channels zero-weighted) -- the split engine.PackedWeights.refresh packs (unet.py:770-788, 545-550)."""
    from ccdm_b200.engine import feat_fold_split
    torch.manual_seed(0)
    cout, H, W = 32, 6, 10
    h, ft = torch.randn(2, c_h, H, W, dtype=torch.float64) * 1.5 + 0.2, torch.randn(2, f, H, W, dtype=torch.float64) * 0.8 - 0.1
    g, be = 1 + 0.1 * torch.randn(c_h + f, dtype=torch.float64), 0.1 * torch.randn(c_h + f, dtype=torch.float64)
    w = torch.randn(cout, c_h + f, 3, 3, dtype=torch.float64) / 30
    want = torch.nn.functional.conv2d(torch.nn.functional.silu(torch.nn.functional.group_norm(torch.cat([h, ft], 1), 32, g, be, 1e-5)), w, padding=1)
    cpg, n_f, f16 = feat_fold_split(c_h, f)
    assert cpg == (c_h + f) // 32 and (c_h + n_f) % cpg == 0 and 0 <= n_f < cpg and f16 % 16 == 0 and f16 >= n_f

    def gn_part(x, off, gam, bet):
        """GroupNorm of channel range [off, off + C) of the concatenation with the statistics of the channels it SEES (what the
        kernels do with ccdm_op::gn_cpg / gn_off): complete groups are exact, cut groups are garbage (zero-weighted)."""
        C = x.shape[1]
        out = torch.zeros_like(x)
        for c in range(C):
            g0 = ((c + off) // cpg) * cpg - off
            j0, j1 = max(g0, 0), min(g0 + cpg, C)
            n = cpg * H * W
            mean = x[:, j0:j1].sum(dim=(1, 2, 3)) / n
            var = (x[:, j0:j1] ** 2).sum(dim=(1, 2, 3)) / n - mean ** 2
            out[:, c] = (x[:, c] - mean[:, None, None]) / torch.sqrt(var.clamp(min=0)[:, None, None] + 1e-5) * gam[c] + bet[c]
        return out

    cs = c_h + f16
    xs = torch.cat([h, ft[:, :f16]], 1)
    w1s = w[:, :cs].clone(); w1s[:, c_h + n_f:] = 0
    w1m = w[:, c_h:].clone(); w1m[:, :n_f] = 0
    step = torch.nn.functional.conv2d(torch.nn.functional.silu(gn_part(xs, 0, g[:cs], be[:cs])), w1s, padding=1)
    chain = torch.nn.functional.conv2d(torch.nn.functional.silu(gn_part(ft, c_h, g[c_h:], be[c_h:])), w1m, padding=1)
    assert float((step + chain - want).abs().max()) < 1e-10


def test_feature_fold_is_planned_for_the_dino_block():
    """The Cityscapes program runs input_blocks[10] in its folded form in every precision mode, on the tensor-core kernel in the
    tensor-core modes, with two per-chain ops beside the step."""
    from ccdm_b200 import _lib
    m = _build(3, 256, 512, 20, True, None)
    for prec in ("exact", "bf16", "fp32"):
        eng = m.unet.engine(prec, dry_run=True)
        eng.weights.refresh()
        prog = eng.program(2, 256, 512)
        prog.bind(8)
        assert len(prog._pre_array) == 2 and not prog.off_tc
        pre_conv, pre_skip = prog._pre_array
        assert (pre_conv.C0, pre_conv.ksize, pre_conv.gn, pre_conv.gn_cpg, pre_conv.gn_off) == (384, 3, 1, 14, 64)
        assert (pre_skip.C0, pre_skip.ksize, pre_skip.gn) == (384, 1, 0)
        folded = [o for o in prog._op_array if o.kind == _lib.OP_CONV and o.gn_cpg > 0]
        assert len(folded) == 1 and (folded[0].C0, folded[0].C1, folded[0].gn_cpg, folded[0].gn_off) == (64, 16, 14, 0)
        assert not [o for o in prog._op_array if o.kind == _lib.OP_CONV and (o.C0 + o.C1 > 256 or o.S0 + o.S1 > 256)]  # no 448-channel op is left


def test_docs_quote_the_current_abi_version():
    """INTEGRATION.md's binding snippet asserts the ABI version: it must be the one the header defines and _lib.py checks."""
    import re
    from ccdm_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "ccdm_b200.h")).read()
    doc = open(os.path.join(root, "INTEGRATION.md")).read()
    v_hdr = int(re.search(r"#define CCDM_ABI_VERSION (\d+)", hdr).group(1))
    v_doc = [int(v) for v in re.findall(r"ccdm_abi_version\(\) == (\d+)", doc)]
    assert v_hdr == _lib.ABI_VERSION and v_doc and all(v == v_hdr for v in v_doc)


def test_vit_qkv_row_permutation_matches_the_attention_kernels_order():
    """The encoder packs a ViT block's qkv rows [which][head][d] (Attention.forward: reshape(B, N, 3, heads, d)) into the
    attention kernel's per-head order [head][which][d] (QKVAttentionLegacy, unet.py:343-360): row r of the packed weight is
    row perm[r] of the original."""
    D, heads, hd = 384, 6, 64
    perm = torch.arange(3 * D).reshape(3, heads, hd).permute(1, 0, 2).reshape(-1)
    for h in (0, 3, 5):
        for which in (0, 1, 2):
            for d in (0, 17, 63):
                assert int(perm[h * 3 * hd + which * hd + d]) == which * D + h * hd + d
    assert sorted(perm.tolist()) == list(range(3 * D))
